# builds libb200ssm variants with debug switches and times the stage-0 forward (results are WRONG by construction: timing only)
set -e
cd /root/repo
base="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr"
L=medical_image_classification_b200/lib
cp $L/libb200ssm.so /tmp/lib_orig.so
for v in "" "-DB200_DBG_NO_EX2" "-DB200_DBG_NO_BC" "-DB200_DBG_NO_EX2 -DB200_DBG_NO_BC" "-DB200_DBG_NO_SHFL" "-DB200_DBG_NO_SHFL -DB200_DBG_NO_EX2"; do
  nvcc $base $v -c medical_image_classification_b200/csrc/sscan2.cu -o /tmp/sscan2_v.o
  nvcc -shared -o $L/libb200ssm.so $L/api.o $L/cross.o $L/dwconv.o $L/glue.o $L/lngate.o $L/sscan.o /tmp/sscan2_v.o $L/ssd.o -lcudart
  echo "== variant [$v]"
  python tools/prof_sscan.py 0 64 3 2>&1 | tail -1
done
cp /tmp/lib_orig.so $L/libb200ssm.so
