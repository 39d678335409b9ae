// stubs.cu -- entry points declared in include/b200_ssm.h whose kernels are not built yet.
// Each returns -2 with a message; none of them silently succeeds.
#include "common.cuh"
using namespace b200;
#define NOT_YET(name) do { set_error(name ": not implemented in this build"); return -2; } while (0)
extern "C" size_t b200_ssd_workspace_bytes(int32_t, int32_t, int32_t, int32_t, int32_t, int32_t, int32_t) { return 0; }
extern "C" size_t b200_ssd_bwd_scratch_bytes(int32_t, int32_t, int32_t, int32_t, int32_t, int32_t, int32_t) { return 0; }
extern "C" int b200_ssd_fwd(const b200_ssd_fwd_params*, b200_stream_t) { NOT_YET("b200_ssd_fwd"); }
extern "C" int b200_ssd_bwd(const b200_ssd_bwd_params*, b200_stream_t) { NOT_YET("b200_ssd_bwd"); }
extern "C" int b200_rmsnorm_gated_fwd(const float*, const float*, const float*, float*, float*, int64_t, int32_t, float, b200_stream_t) { NOT_YET("b200_rmsnorm_gated_fwd"); }
extern "C" int b200_rmsnorm_gated_bwd(const float*, const float*, const float*, const float*, const float*, float*, float*, float*, int32_t, int64_t, int32_t, b200_stream_t) { NOT_YET("b200_rmsnorm_gated_bwd"); }
