// linear_tc.cu -- tcgen05 / TMEM / TMA GEMM engine (placeholder until the tensor-core
// kernels land: falls through to the fp32 engine so mode 1 stays functional).
#include "common.cuh"
namespace mtb {
int linear_fwd_simt(const mtb_linear_desc* d, int n, cudaStream_t st);
int linear_bwd_simt(const mtb_linear_bwd_desc* d, int n, cudaStream_t st);
int linear_fwd_tc(const mtb_linear_desc* d, int n, cudaStream_t st) { return linear_fwd_simt(d, n, st); }
int linear_bwd_tc(const mtb_linear_bwd_desc* d, int n, cudaStream_t st) { return linear_bwd_simt(d, n, st); }
}
