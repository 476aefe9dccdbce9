// api.cu -- C-ABI entry points of libmultb200.so (see include/multb200.h).
#include "common.cuh"
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>

namespace mtb {

static thread_local char g_err[512] = "";
int g_gemm_mode = 0;
int g_attn_mode = -1;
int g_pdl = getenv("MTB_PDL") ? atoi(getenv("MTB_PDL")) : 1;   // programmatic dependent launch (common.cuh)   // -1 = follow the GEMM engine (tensor-core attention in tensor-core mode)
unsigned long long g_launches = 0;
void note_launch() { ++g_launches; }
unsigned long long launches() { return g_launches; }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

int linear_fwd_simt(const mtb_linear_desc* d, int n, cudaStream_t st);
int linear_bwd_simt(const mtb_linear_bwd_desc* d, int n, cudaStream_t st);
int linear_fwd_tc(const mtb_linear_desc* d, int n, cudaStream_t st);
int linear_bwd_tc(const mtb_linear_bwd_desc* d, int n, cudaStream_t st);
int attn_fwd_simt(const mtb_attn_desc* d, int n, cudaStream_t st);
int attn_bwd_simt(const mtb_attn_bwd_desc* d, int n, cudaStream_t st);
int attn_fwd_tc(const mtb_attn_desc* d, int n, cudaStream_t st);
int preload_elementwise();
int preload_layernorm();
int preload_linear_simt();
int preload_linear_tc();
int preload_attention_simt();
int preload_attention_tc();
int attn_bwd_tc(const mtb_attn_bwd_desc* d, int n, cudaStream_t st);
int adam_step(const mtb_adam_desc* d, cudaStream_t st);
int preload_optim();

}  // namespace mtb

extern "C" {

int mtb_abi_version(void) { return MTB_ABI_VERSION; }
const char* mtb_last_error(void) { return mtb::g_err; }
int mtb_sm_count(void) { return mtb::sm_count(); }
int mtb_set_gemm_mode(int mode) {
  const int prev = mtb::g_gemm_mode;
  mtb::g_gemm_mode = mode <= 0 ? 0 : (mode >= 2 ? 2 : 1);
  return prev;
}
int mtb_get_gemm_mode(void) { return mtb::g_gemm_mode; }
int mtb_preload(void) {
  const int bad = mtb::preload_elementwise() + mtb::preload_layernorm() + mtb::preload_linear_simt() + mtb::preload_linear_tc() +
                  mtb::preload_attention_simt() + mtb::preload_attention_tc() + mtb::preload_optim();
  MTB_CHECK(bad == 0, "preload: %d kernels failed to load (%s)", bad, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
int mtb_set_attn_mode(int mode) {
  const int prev = mtb::g_attn_mode;
  mtb::g_attn_mode = mode < 0 ? -1 : (mode ? 1 : 0);
  return prev;
}
int mtb_get_attn_mode(void) { return mtb::g_attn_mode; }
uint64_t mtb_launch_count(void) { return mtb::launches(); }

int mtb_linear_fwd(const mtb_linear_desc* d, int n, void* stream) {
  MTB_CHECK(n >= 1 && n <= MTB_MAX_GROUP, "linear_fwd: group size %d out of range", n);
  for (int i = 0; i < n; ++i) {
    MTB_CHECK(d[i].X && d[i].W && d[i].Y, "linear_fwd: null operand in problem %d", i);
    MTB_CHECK(d[i].M >= 0 && d[i].N >= 0 && d[i].K >= 0, "linear_fwd: negative size in problem %d", i);
    MTB_CHECK(d[i].act == 0 || d[i].act == 1, "linear_fwd: unknown activation %d", d[i].act);
  }
  if (mtb::g_gemm_mode >= 1) return mtb::linear_fwd_tc(d, n, (cudaStream_t)stream);
  return mtb::linear_fwd_simt(d, n, (cudaStream_t)stream);
}

int mtb_linear_bwd(const mtb_linear_bwd_desc* d, int n, void* stream) {
  MTB_CHECK(n >= 1 && n <= MTB_MAX_GROUP, "linear_bwd: group size %d out of range", n);
  for (int i = 0; i < n; ++i) {
    MTB_CHECK(d[i].dY && d[i].W, "linear_bwd: null operand in problem %d", i);
    MTB_CHECK(!d[i].db || d[i].dW || mtb::g_gemm_mode >= 1,
              "linear_bwd: the fp32 engine computes the bias gradient inside the weight-gradient GEMM (problem %d)", i);
    MTB_CHECK(!d[i].dW || d[i].X, "linear_bwd: weight gradient needs the forward input (problem %d)", i);
    MTB_CHECK(d[i].act == 0 || d[i].Yact, "linear_bwd: act=1 needs the forward output (problem %d)", i);
  }
  if (mtb::g_gemm_mode >= 1) return mtb::linear_bwd_tc(d, n, (cudaStream_t)stream);
  return mtb::linear_bwd_simt(d, n, (cudaStream_t)stream);
}

int mtb_attn_fwd(const mtb_attn_desc* d, int n, void* stream) {
  MTB_CHECK(n >= 1 && n <= MTB_MAX_GROUP, "attn_fwd: group size %d out of range", n);
  for (int i = 0; i < n; ++i)
    MTB_CHECK(d[i].q && d[i].k && d[i].v && d[i].o && d[i].Lq > 0 && d[i].Lk > 0 && d[i].hd > 0,
              "attn_fwd: bad problem %d", i);
  if (mtb::g_attn_mode == 1 || (mtb::g_attn_mode < 0 && mtb::g_gemm_mode >= 1)) return mtb::attn_fwd_tc(d, n, (cudaStream_t)stream);
  return mtb::attn_fwd_simt(d, n, (cudaStream_t)stream);
}

int mtb_attn_bwd(const mtb_attn_bwd_desc* d, int n, void* stream) {
  MTB_CHECK(n >= 1 && n <= MTB_MAX_GROUP, "attn_bwd: group size %d out of range", n);
  for (int i = 0; i < n; ++i)
    MTB_CHECK(d[i].q && d[i].k && d[i].v && d[i].o && d[i].d_o && d[i].lse && d[i].delta && d[i].dq && d[i].dk && d[i].dv,
              "attn_bwd: null operand in problem %d", i);
  if (mtb::g_attn_mode == 1 || (mtb::g_attn_mode < 0 && mtb::g_gemm_mode >= 1)) return mtb::attn_bwd_tc(d, n, (cudaStream_t)stream);
  return mtb::attn_bwd_simt(d, n, (cudaStream_t)stream);
}

static int run_ops_impl(const mtb_op* ops, int n_ops, void* stream, void* side_stream, cudaEvent_t ev_fork, cudaEvent_t ev_join) {
  cudaStream_t st = (cudaStream_t)stream, s2 = (cudaStream_t)side_stream;
  bool forked = false;
  for (int i = 0; i < n_ops; ++i) {
    const mtb_op& o = ops[i];
    void* tgt = stream;
    if (s2 && o.side) {
      MTB_CUDA(cudaEventRecord(ev_fork, st));
      MTB_CUDA(cudaStreamWaitEvent(s2, ev_fork, 0));
      tgt = side_stream;
      forked = true;
    }
    int rc;
    switch (o.kind) {
      case MTB_OP_EMBED_FWD: rc = mtb_embed_fwd((const mtb_embed_desc*)o.descs, o.n, tgt); break;
      case MTB_OP_EMBED_BWD: rc = mtb_embed_bwd((const mtb_embed_desc*)o.descs, o.n, tgt); break;
      case MTB_OP_ADDN: rc = mtb_addn((const mtb_addn_desc*)o.descs, o.n, tgt); break;
      case MTB_OP_RESLN_FWD: rc = mtb_resln_fwd((const mtb_resln_desc*)o.descs, o.n, tgt); break;
      case MTB_OP_RESLN_BWD: rc = mtb_resln_bwd((const mtb_resln_bwd_desc*)o.descs, o.n, tgt); break;
      case MTB_OP_LINEAR_FWD: rc = mtb_linear_fwd((const mtb_linear_desc*)o.descs, o.n, tgt); break;
      case MTB_OP_LINEAR_BWD: rc = mtb_linear_bwd((const mtb_linear_bwd_desc*)o.descs, o.n, tgt); break;
      case MTB_OP_ATTN_FWD: rc = mtb_attn_fwd((const mtb_attn_desc*)o.descs, o.n, tgt); break;
      case MTB_OP_ATTN_BWD: rc = mtb_attn_bwd((const mtb_attn_bwd_desc*)o.descs, o.n, tgt); break;
      default: MTB_CHECK(false, "run_ops: unknown op kind %d at position %d", o.kind, i);
    }
    if (rc != 0) return (i + 1) * 16 + (rc < 0 ? -rc : rc);   // position of the failing op (message already set)
  }
  if (forked) {
    MTB_CUDA(cudaEventRecord(ev_join, s2));
    MTB_CUDA(cudaStreamWaitEvent(st, ev_join, 0));
  }
  return 0;
}

int mtb_run_ops(const mtb_op* ops, int n_ops, void* stream, void* side_stream) {
  MTB_CHECK(ops != nullptr || n_ops == 0, "run_ops: null op list");
  static cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  if (side_stream && !ev_fork) {
    MTB_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    MTB_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
  }
  return run_ops_impl(ops, n_ops, stream, side_stream, ev_fork, ev_join);
}

/* ---- CUDA graphs of op lists ------------------------------------------------------------------------------------
 * Every address a stage batch touches is persistent (plan executor regions, parameters, the device-side dropout
 * counter), so an op list that comes back -- the same stage composition in a later step -- can be replayed as ONE
 * graph launch instead of ~20-60 kernel launches: the host cost of a stage drops from ~0.2 ms to ~10 us and the
 * GPU-side launch gaps go with it.  Side-stream ops become a parallel branch of the graph. */
struct MtbGraph {
  cudaGraph_t graph;
  cudaGraphExec_t exec;
  unsigned long long launches;
};

int mtb_graph_capture(const mtb_op* ops, int n_ops, int use_side, void** out) {
  MTB_CHECK(ops != nullptr && n_ops > 0 && out != nullptr, "graph_capture: empty op list");
  static cudaStream_t cap = nullptr, cap2 = nullptr;
  static cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  if (!cap) {
    MTB_CUDA(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
    MTB_CUDA(cudaStreamCreateWithFlags(&cap2, cudaStreamNonBlocking));
    MTB_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    MTB_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
  }
  *out = nullptr;
  for (int attempt = 0; attempt < 2; ++attempt) {
    const int pdl0 = mtb::g_pdl;
    if (attempt == 1) mtb::g_pdl = 0;            // second try without programmatic dependent launch edges
    const unsigned long long l0 = mtb::g_launches;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    cudaError_t e = cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal);
    int rc = -1;
    if (e == cudaSuccess) {
      rc = run_ops_impl(ops, n_ops, cap, use_side ? cap2 : nullptr, ev_fork, ev_join);
      e = cudaStreamEndCapture(cap, &graph);
    }
    const unsigned long long captured = mtb::g_launches - l0;
    mtb::g_launches = l0;                        // nothing ran
    mtb::g_pdl = pdl0;
    if (rc == 0 && e == cudaSuccess && graph != nullptr) e = cudaGraphInstantiate(&exec, graph, 0);
    if (rc == 0 && e == cudaSuccess && exec != nullptr) {
      MtbGraph* g = new MtbGraph{graph, exec, captured};
      *out = g;
      return 0;
    }
    if (graph) cudaGraphDestroy(graph);
    (void)cudaGetLastError();
    if (attempt == 1) {
      if (rc == 0) mtb::set_error("graph_capture: %s", cudaGetErrorString(e));
      return -3;
    }
  }
  return -3;
}

int mtb_graph_launch(void* handle, void* stream) {
  MTB_CHECK(handle != nullptr, "graph_launch: null handle");
  MtbGraph* g = (MtbGraph*)handle;
  MTB_CUDA(cudaGraphLaunch(g->exec, (cudaStream_t)stream));
  mtb::g_launches += g->launches;               // kernels the replay runs (bench bookkeeping)
  return 0;
}

int mtb_graph_destroy(void* handle) {
  if (handle == nullptr) return 0;
  MtbGraph* g = (MtbGraph*)handle;
  cudaGraphExecDestroy(g->exec);
  cudaGraphDestroy(g->graph);
  delete g;
  return 0;
}

int mtb_adam_step(const mtb_adam_desc* d, void* stream) {
  MTB_CHECK(d != nullptr, "adam_step: null descriptor");
  return mtb::adam_step(d, (cudaStream_t)stream);
}

}  // extern "C"
