# A/B of ssd.cu build variants on one box: VARIANTS="flags;flags;..." tools/ssd_variants.sh [prof_ssd args]
set -e
cd /root/repo
base="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr"
L=medical_image_classification_b200/lib
cp $L/libb200ssm.so /tmp/lib_orig.so
VARIANTS=${VARIANTS:-";-DB200_SSD_TC_CTAS=2"}
IFS=';' read -ra VS <<< "$VARIANTS"
for v in "${VS[@]}"; do
  nvcc $base $v -c medical_image_classification_b200/csrc/ssd.cu -o /tmp/ssd_v.o 2>/dev/null
  nvcc -shared -o $L/libb200ssm.so $L/api.o $L/cross.o $L/dwconv.o $L/glue.o $L/lngate.o $L/sscan.o $L/sscan2.o /tmp/ssd_v.o -lcudart 2>/dev/null
  echo "== variant [$v]"
  python tools/prof_ssd.py ${@:-0 64 1 3} 2>&1 | tail -1
done
cp /tmp/lib_orig.so $L/libb200ssm.so
