// Dispatch of the tensor-core attention kernels (attention_tc.cuh: TF32 mma on fp32 tiles; attention_tc16.cuh: bf16 mma
// with cp.async / ldmatrix) by storage type, head width and padded length.
#include "../../../include/morec_b200.h"
#include "attention_tc16.cuh"

namespace morec {

template <typename T>
static int tc_by_shape(const TcAttnParams& p, bool bwd, cudaStream_t stream) {
    const bool small = p.seqlen <= 32;
    if (p.head_dim == 256) return tc_launch<T, 256, 32>(p, bwd, stream);
    if (p.head_dim == 64) return small ? tc_launch<T, 64, 32>(p, bwd, stream) : tc_launch<T, 64, 64>(p, bwd, stream);
    return small ? tc_launch<T, 32, 32>(p, bwd, stream) : tc_launch<T, 32, 64>(p, bwd, stream);
}

template <typename T>
static int tc16_by_shape(const TcAttnParams& p, bool bwd, cudaStream_t stream) {
    const bool small = p.seqlen <= 32;
    if (p.head_dim == 64) return small ? tc16_launch<T, 64, 32>(p, bwd, stream) : tc16_launch<T, 64, 64>(p, bwd, stream);
    return small ? tc16_launch<T, 32, 32>(p, bwd, stream) : tc16_launch<T, 32, 64>(p, bwd, stream);
}

int tc_attn_dispatch(const TcAttnParams& p, bool bwd, int dtype, cudaStream_t stream) {
    MOREC_CHECK_ARG(tc_attn_eligible(dtype, p.seqlen, p.head_dim, p.ld, p.ld_o), "attn_tc: unsupported shape");
    MOREC_CHECK_ARG(!p.mask || p.n_mask > 0, "attn_tc: mask needs n_mask > 0");
    if (p.n_seq <= 0) return MOREC_OK;
    if (tc16_eligible(p, dtype, bwd))                                           // 16-bit-native (m16n8k16) kernels
        return dtype == 3 ? tc16_by_shape<__half>(p, bwd, stream) : tc16_by_shape<__nv_bfloat16>(p, bwd, stream);
    MOREC_DISPATCH_T(dtype, return tc_by_shape<T>(p, bwd, stream));
    return MOREC_OK;
}

}  // namespace morec
