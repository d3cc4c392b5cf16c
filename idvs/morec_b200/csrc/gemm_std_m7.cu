// tcgen05 GEMM kernels specialised on standard epilogue mode 7 (see MOREC_EPI_* in include/morec_b200.h)
#include "gemm_std_epi.cuh"
MOREC_DEFINE_STD_GEMM(7)
