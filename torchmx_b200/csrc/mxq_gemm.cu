// K3: block-scaled MX GEMM on tcgen05 (kind::mxf8f6f4, E8M0 scale factors in TMEM).  Placeholder
// until the tensor-core kernel lands: reports "unsupported shape" so the host takes the
// dequantize path (which is what the reference itself does, torchmx/ops.py:29-41).
#include <cstdio>

#include "mxq_common.cuh"

namespace mxq {

int launch_gemm(const mxq_gemm_args_t* a, int sm_count, cudaStream_t stream, char* msg, size_t msg_len) {
    (void)a; (void)sm_count; (void)stream;
    snprintf(msg, msg_len, "tensor-core path not built yet");
    return MXQ_ERR_UNSUPPORTED_SHAPE;
}

}  // namespace mxq
