// Context management for libnbw: device selection, error string, scratch arena.
#include "nbw_common.cuh"

#include <new>

extern "C" int nbw_version(void) { return 100; }

extern "C" int nbw_create(nbw_ctx** out, int device) {
  if (!out) return NBW_ERR_INVALID;
  *out = nullptr;
  nbw_ctx* ctx = new (std::nothrow) nbw_ctx();
  if (!ctx) return NBW_ERR_INVALID;
  memset(ctx, 0, sizeof(*ctx));
  ctx->device = device;
  cudaError_t e = cudaSetDevice(device);
  cudaDeviceProp prop;
  if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) {
    fprintf(stderr, "nbw_create: %s\n", cudaGetErrorString(e));
    delete ctx;
    return NBW_ERR_CUDA;
  }
  if (prop.major != 10) {
    // sm_100a cubins only: there is deliberately no fallback path
    fprintf(stderr, "nbw_create: device %d is sm_%d%d; libnbw is built for sm_100a (B200) only\n", device,
            prop.major, prop.minor);
    delete ctx;
    return NBW_ERR_UNSUPPORTED;
  }
  ctx->sm_count = prop.multiProcessorCount;
  *out = ctx;
  return NBW_OK;
}

extern "C" int nbw_destroy(nbw_ctx* ctx) {
  if (!ctx) return NBW_OK;
  if (ctx->scratch) cudaFree(ctx->scratch);
  if (ctx->pinned_small) cudaFreeHost(ctx->pinned_small);
  if (ctx->topk_blocks) cudaFree(ctx->topk_blocks);
  if (ctx->bin_order) cudaFree(ctx->bin_order);
  if (ctx->piece_cands) cudaFree(ctx->piece_cands);
  if (ctx->pipe_ready) {
    cudaStreamDestroy(ctx->s_in);
    cudaStreamDestroy(ctx->s_out);
    for (int i = 0; i < 2; ++i) {
      cudaEventDestroy(ctx->ev_ready[i]);
      cudaEventDestroy(ctx->ev_free[i]);
      cudaEventDestroy(ctx->ev_done[i]);
      if (ctx->stage[i]) cudaFree(ctx->stage[i]);
    }
    cudaEventDestroy(ctx->ev_out);
  }
  delete ctx;
  return NBW_OK;
}

extern "C" int nbw_wait_host_scores(nbw_ctx* ctx) {
  if (!ctx) return NBW_ERR_INVALID;
  NBW_CUDA(ctx, nbw_await_host_scores(ctx));
  return NBW_OK;
}

// NVTX ranges for the host layer (fit / score state / broadcast / propose_location), so that a timeline shows the
// BO iteration structure around the kernels; the C entry points open their own ranges (NbwRange).
extern "C" int nbw_range_push(const char* name) { return nvtxRangePushA(name ? name : "nbw"); }
extern "C" int nbw_range_pop(void) { return nvtxRangePop(); }

extern "C" int nbw_profile_events(nbw_ctx* ctx, void* ev_start, void* ev_stop) {
  if (!ctx) return NBW_ERR_INVALID;
  ctx->prof_ev[0] = static_cast<cudaEvent_t>(ev_start);
  ctx->prof_ev[1] = static_cast<cudaEvent_t>(ev_stop);
  return NBW_OK;
}

extern "C" const char* nbw_last_error(const nbw_ctx* ctx) { return ctx ? ctx->err : "null context"; }

extern "C" uint64_t nbw_launch_count(const nbw_ctx* ctx) { return ctx ? ctx->launches : 0; }

int nbw_scratch_ensure(nbw_ctx* ctx, size_t bytes, cudaStream_t stream) {
  if (bytes <= ctx->scratch_bytes) return NBW_OK;
  size_t want = ctx->scratch_bytes ? ctx->scratch_bytes : (size_t)1 << 20;
  while (want < bytes) want *= 2;
  if (ctx->scratch) {
    NBW_CUDA(ctx, cudaStreamSynchronize(stream));
    NBW_CUDA(ctx, cudaDeviceSynchronize());
    NBW_CUDA(ctx, cudaFree(ctx->scratch));
    ctx->scratch = nullptr;
    ctx->scratch_bytes = 0;
  }
  NBW_CUDA(ctx, cudaMalloc(&ctx->scratch, want));
  ctx->scratch_bytes = want;
  return NBW_OK;
}

extern "C" int nbw_reserve(nbw_ctx* ctx, uint64_t scratch_bytes) {
  if (!ctx) return NBW_ERR_INVALID;
  return nbw_scratch_ensure(ctx, (size_t)scratch_bytes, 0);
}
