# A/B of sscan2.cu build variants on one box (timing; the B200_DBG_* switches REMOVE work: their results are wrong by construction)
#   VARIANTS="flags;flags;..." tools/dbg_variants.sh [prof_sscan args]
set -e
cd /root/repo
base="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr"
L=medical_image_classification_b200/lib
cp $L/libb200ssm.so /tmp/lib_orig.so
VARIANTS=${VARIANTS:-";-DB200_DBG_NO_EX2;-DB200_DBG_NO_BC;-DB200_DBG_NO_EX2 -DB200_DBG_NO_BC;-DB200_DBG_NO_SHFL;-DB200_DBG_NO_SHFL -DB200_DBG_NO_EX2"}
IFS=';' read -ra VS <<< "$VARIANTS"
for v in "${VS[@]}"; do
  nvcc $base $v -c medical_image_classification_b200/csrc/sscan2.cu -o /tmp/sscan2_v.o 2>/dev/null
  nvcc -shared -o $L/libb200ssm.so $L/api.o $L/cross.o $L/dwconv.o $L/glue.o $L/lngate.o $L/sscan.o /tmp/sscan2_v.o $L/ssd.o -lcudart 2>/dev/null
  echo "== variant [$v]"
  python tools/prof_sscan.py ${@:-0 64 3} 2>&1 | tail -1
done
cp /tmp/lib_orig.so $L/libb200ssm.so
