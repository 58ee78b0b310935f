// orbp_*: the streaming front end of a multi-camera rig behind the C ABI (include/orb_b200.h) — what Tracking does
// per rig-frame in the reference (Frame::Frame runs one ORBextractor per camera, src/Frame.cc:148-346 / :182-185;
// MonocularInitialization matches consecutive frames of camera 1 with ORBmatcher::SearchForInitialization,
// src/Tracking.cc:870-871), batched over `rig_frames` rig-frames per step and software-pipelined over four CUDA
// streams so that a C++ host gets the overlapped path, not only the Python harness:
//
//   copy-in   H2D of the step's frames                         (DMA engine)
//   compute   extractor of every camera, ONE stream: the extractor kernels fill the GPU on their own
//   match     SearchForInitialization of camera 0 on a high-priority side stream as soon as camera 0 is
//             extracted: its ordered resolve is a serial chain per pair that leaves the SMs idle, so it runs
//             underneath the extraction of the other cameras
//   copy-out  D2H of keypoints / descriptors / counts / matches into pinned host memory (the other DMA engine)
//
// Device image buffers and output buffers are `depth` steps deep: the copies of step k+1 overlap the kernels of
// step k while the consumer still reads step k-1.  orbp_submit never blocks the host; orbp_wait blocks for one step.
// Built on the library's own entry points (orbx_*, orbm_*); there is no CPU fallback.
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/orb_b200.h"
#include "device_guard.h"

namespace {
thread_local std::string g_pcreate_error;

struct Slot {
  orbx_keypoint* d_kps[ORBP_MAX_CAMS] = {};
  uint8_t* d_desc[ORBP_MAX_CAMS] = {};
  int32_t* d_counts[ORBP_MAX_CAMS] = {};
  int32_t* d_m12 = nullptr;
  int32_t* d_nm = nullptr;
  orbx_keypoint* h_kps[ORBP_MAX_CAMS] = {};
  uint8_t* h_desc[ORBP_MAX_CAMS] = {};
  int32_t* h_counts[ORBP_MAX_CAMS] = {};
  int32_t* h_m12 = nullptr;
  int32_t* h_nm = nullptr;
  uint8_t* d_img[ORBP_MAX_CAMS] = {};
  cudaEvent_t ev_in[ORBP_MAX_CAMS] = {}, ev_done[ORBP_MAX_CAMS] = {};
  cudaEvent_t ev_cam0 = nullptr, ev_match = nullptr, done = nullptr;
  bool used = false;
};
}  // namespace

struct orbp_pipeline {
  orbp_config cfg;
  int device = 0;
  int pitch = 0;  // device row pitch of the frames (width rounded up to 16: level 0 is then read in place)
  int caps[ORBP_MAX_CAMS] = {};
  cudaStream_t s_in = nullptr, s_compute = nullptr, s_match = nullptr, s_out = nullptr;
  orbx_extractor* ex[ORBP_MAX_CAMS] = {};
  orbm_matcher* matcher = nullptr;
  std::vector<Slot> slots;
  long long n_submitted = 0;
  std::string err;
  bool check(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return true;
    err = std::string(what) + ": " + cudaGetErrorString(e);
    return false;
  }
};

extern "C" {
#pragma GCC visibility push(default)

const char* orbp_last_error(const orbp_pipeline* p) { return p ? p->err.c_str() : g_pcreate_error.c_str(); }

void orbp_destroy(orbp_pipeline* p) {
  if (!p) return;
  OrbDeviceGuard guard(p->device);
  for (cudaStream_t s : {p->s_in, p->s_compute, p->s_match, p->s_out})
    if (s) cudaStreamSynchronize(s);
  for (int c = 0; c < ORBP_MAX_CAMS; ++c) orbx_destroy(p->ex[c]);
  orbm_destroy(p->matcher);
  for (Slot& s : p->slots) {
    for (int c = 0; c < ORBP_MAX_CAMS; ++c) {
      cudaFree(s.d_kps[c]); cudaFree(s.d_desc[c]); cudaFree(s.d_counts[c]); cudaFree(s.d_img[c]);
      if (s.h_kps[c]) cudaFreeHost(s.h_kps[c]);
      if (s.h_desc[c]) cudaFreeHost(s.h_desc[c]);
      if (s.h_counts[c]) cudaFreeHost(s.h_counts[c]);
      if (s.ev_in[c]) cudaEventDestroy(s.ev_in[c]);
      if (s.ev_done[c]) cudaEventDestroy(s.ev_done[c]);
    }
    cudaFree(s.d_m12); cudaFree(s.d_nm);
    if (s.h_m12) cudaFreeHost(s.h_m12);
    if (s.h_nm) cudaFreeHost(s.h_nm);
    for (cudaEvent_t e : {s.ev_cam0, s.ev_match, s.done})
      if (e) cudaEventDestroy(e);
  }
  for (cudaStream_t s : {p->s_in, p->s_compute, p->s_match, p->s_out})
    if (s) cudaStreamDestroy(s);
  delete p;
}

int orbp_create(const orbp_config* cfg, orbp_pipeline** out) {
  if (!cfg || !out) { g_pcreate_error = "null argument"; return ORBX_E_INVALID; }
  *out = nullptr;
  if (cfg->n_cams < 1 || cfg->n_cams > ORBP_MAX_CAMS || cfg->rig_frames < 1 || cfg->width < 1 || cfg->height < 1) {
    g_pcreate_error = "invalid pipeline configuration";
    return ORBX_E_INVALID;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    g_pcreate_error = "no CUDA device (this library has no CPU fallback)";
    return ORBX_E_CUDA;
  }
  orbp_pipeline* p = new orbp_pipeline();
  p->cfg = *cfg;
  if (p->cfg.depth < 2) p->cfg.depth = 3;
  auto fail = [&](int code) {
    g_pcreate_error = p->err;
    orbp_destroy(p);
    return code;
  };
  if (cfg->device >= ndev) { p->err = "no such CUDA device"; return fail(ORBX_E_CUDA); }
  if (cfg->device >= 0) p->device = cfg->device; else cudaGetDevice(&p->device);
  OrbDeviceGuard guard(p->device);
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);  // hi = numerically lowest = highest priority
  if (!p->check(cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking), "stream") ||
      !p->check(cudaStreamCreateWithFlags(&p->s_compute, cudaStreamNonBlocking), "stream") ||
      !p->check(cudaStreamCreateWithPriority(&p->s_match, cudaStreamNonBlocking, hi), "stream") ||
      !p->check(cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking), "stream"))
    return fail(ORBX_E_CUDA);
  const int F = cfg->rig_frames;
  p->pitch = (cfg->width + 15) & ~15;
  for (int c = 0; c < cfg->n_cams; ++c) {
    orbx_config xc = {cfg->nfeatures[c], cfg->scale_factor, cfg->nlevels, cfg->ini_th_fast, cfg->min_th_fast, cfg->width,
                      cfg->height, F, p->device};
    if (orbx_create(&xc, &p->ex[c]) != ORBX_OK) { p->err = orbx_last_error(nullptr); return fail(ORBX_E_CUDA); }
    if (orbx_set_stream(p->ex[c], p->s_compute) != ORBX_OK) { p->err = orbx_last_error(p->ex[c]); return fail(ORBX_E_CUDA); }
    p->caps[c] = orbx_max_keypoints(p->ex[c]);
  }
  if (orbm_create(p->device, &p->matcher) != ORBX_OK) { p->err = orbm_last_error(nullptr); return fail(ORBX_E_CUDA); }
  if (orbm_set_stream(p->matcher, p->s_match) != ORBX_OK) { p->err = orbm_last_error(p->matcher); return fail(ORBX_E_CUDA); }
  p->slots.resize(p->cfg.depth);
  const bool match = cfg->match && F > 1;
  for (Slot& s : p->slots) {
    bool ok = true;
    for (int c = 0; c < cfg->n_cams && ok; ++c) {
      const size_t nk = (size_t)F * p->caps[c];
      ok = p->check(cudaMalloc((void**)&s.d_kps[c], nk * sizeof(orbx_keypoint)), "cudaMalloc(kps)") &&
           p->check(cudaMalloc((void**)&s.d_desc[c], nk * 32), "cudaMalloc(desc)") &&
           p->check(cudaMalloc((void**)&s.d_counts[c], sizeof(int32_t) * F), "cudaMalloc(counts)") &&
           p->check(cudaMalloc((void**)&s.d_img[c], (size_t)F * p->pitch * cfg->height), "cudaMalloc(frames)") &&
           p->check(cudaHostAlloc((void**)&s.h_kps[c], nk * sizeof(orbx_keypoint), cudaHostAllocDefault), "cudaHostAlloc(kps)") &&
           p->check(cudaHostAlloc((void**)&s.h_desc[c], nk * 32, cudaHostAllocDefault), "cudaHostAlloc(desc)") &&
           p->check(cudaHostAlloc((void**)&s.h_counts[c], sizeof(int32_t) * F, cudaHostAllocDefault), "cudaHostAlloc(counts)") &&
           p->check(cudaEventCreateWithFlags(&s.ev_in[c], cudaEventDisableTiming), "event") &&
           p->check(cudaEventCreateWithFlags(&s.ev_done[c], cudaEventDisableTiming), "event");
    }
    if (ok && match) {
      const size_t nm = (size_t)(F - 1) * p->caps[0];
      ok = p->check(cudaMalloc((void**)&s.d_m12, nm * sizeof(int32_t)), "cudaMalloc(matches)") &&
           p->check(cudaMalloc((void**)&s.d_nm, sizeof(int32_t) * (F - 1)), "cudaMalloc(nmatches)") &&
           p->check(cudaHostAlloc((void**)&s.h_m12, nm * sizeof(int32_t), cudaHostAllocDefault), "cudaHostAlloc(matches)") &&
           p->check(cudaHostAlloc((void**)&s.h_nm, sizeof(int32_t) * (F - 1), cudaHostAllocDefault), "cudaHostAlloc(nmatches)");
    }
    ok = ok && p->check(cudaEventCreateWithFlags(&s.ev_cam0, cudaEventDisableTiming), "event") &&
         p->check(cudaEventCreateWithFlags(&s.ev_match, cudaEventDisableTiming), "event") &&
         p->check(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming), "event");
    if (!ok) return fail(ORBX_E_CUDA);
  }
  *out = p;
  return ORBX_OK;
}

int orbp_capacity(const orbp_pipeline* p, int cam) {
  return (p && cam >= 0 && cam < p->cfg.n_cams) ? p->caps[cam] : ORBX_E_INVALID;
}

long long orbp_submit(orbp_pipeline* p, const uint8_t* const* images, size_t frame_stride, size_t row_stride) {
  if (!p || !images) return ORBX_E_INVALID;
  const orbp_config& cfg = p->cfg;
  const int F = cfg.rig_frames, W = cfg.width, H = cfg.height, P = p->pitch;
  if (row_stride < (size_t)W || frame_stride < row_stride * (size_t)H) { p->err = "invalid strides"; return ORBX_E_INVALID; }
  for (int c = 0; c < cfg.n_cams; ++c)
    if (!images[c]) { p->err = "null image pointer"; return ORBX_E_INVALID; }
  OrbDeviceGuard guard(p->device);
  const long long step = p->n_submitted;
  Slot& s = p->slots[step % cfg.depth];
  const bool match = cfg.match && F > 1;
  // the slot's previous results must have left the device before its buffers (frames, outputs) are reused; that
  // event lies `depth` steps back, so this is not a stall in steady state
  if (s.used) {
    cudaStreamWaitEvent(p->s_in, s.done, 0);
    cudaStreamWaitEvent(p->s_compute, s.done, 0);
  }
  // per-camera events: camera 0 starts computing while camera 1 is still uploading, and its features leave the device
  // while camera 1 computes
  for (int c = 0; c < cfg.n_cams; ++c) {
    cudaError_t e;
    if (frame_stride == row_stride * (size_t)H)
      e = cudaMemcpy2DAsync(s.d_img[c], P, images[c], row_stride, W, (size_t)H * F, cudaMemcpyHostToDevice, p->s_in);
    else {
      e = cudaSuccess;
      for (int f = 0; f < F && e == cudaSuccess; ++f)
        e = cudaMemcpy2DAsync(s.d_img[c] + (size_t)f * P * H, P, images[c] + (size_t)f * frame_stride, row_stride, W, H,
                              cudaMemcpyHostToDevice, p->s_in);
    }
    if (!p->check(e, "H2D frames")) return ORBX_E_CUDA;
    cudaEventRecord(s.ev_in[c], p->s_in);
  }
  for (int c = 0; c < cfg.n_cams; ++c) {
    cudaStreamWaitEvent(p->s_compute, s.ev_in[c], 0);
    if (orbx_extract_batch_device(p->ex[c], s.d_img[c], F, (size_t)P * H, P, s.d_kps[c], s.d_desc[c], s.d_counts[c],
                                  p->caps[c]) != ORBX_OK) {
      p->err = orbx_last_error(p->ex[c]);
      return ORBX_E_CUDA;
    }
    cudaEventRecord(s.ev_done[c], p->s_compute);
    if (c == 0 && match) {
      cudaStreamWaitEvent(p->s_match, s.ev_done[0], 0);
      // pairs (t, t+1) of camera 0: the F2 arrays are the same buffers shifted by one frame; vbPrevMatched = F1's
      // keypoint positions (how Tracking initialises it, src/Tracking.cc:844-846)
      const orbm_bounds b = {0.0f, (float)W, 0.0f, (float)H};
      const int cap = p->caps[0];
      if (orbm_search_for_initialization_device(p->matcher, F - 1, cap, s.d_kps[0], s.d_desc[0], s.d_counts[0], s.d_kps[0] + cap,
                                                s.d_desc[0] + (size_t)cap * 32, s.d_counts[0] + 1, b, nullptr, cfg.window,
                                                cfg.nnratio, cfg.check_ori, s.d_m12, s.d_nm) != ORBX_OK) {
        p->err = orbm_last_error(p->matcher);
        return ORBX_E_CUDA;
      }
      cudaEventRecord(s.ev_match, p->s_match);
    }
  }
  for (int c = 0; c < cfg.n_cams; ++c) {
    const size_t nk = (size_t)F * p->caps[c];
    cudaStreamWaitEvent(p->s_out, s.ev_done[c], 0);
    cudaMemcpyAsync(s.h_kps[c], s.d_kps[c], nk * sizeof(orbx_keypoint), cudaMemcpyDeviceToHost, p->s_out);
    cudaMemcpyAsync(s.h_desc[c], s.d_desc[c], nk * 32, cudaMemcpyDeviceToHost, p->s_out);
    cudaMemcpyAsync(s.h_counts[c], s.d_counts[c], sizeof(int32_t) * F, cudaMemcpyDeviceToHost, p->s_out);
  }
  if (match) {
    cudaStreamWaitEvent(p->s_out, s.ev_match, 0);
    cudaMemcpyAsync(s.h_m12, s.d_m12, (size_t)(F - 1) * p->caps[0] * sizeof(int32_t), cudaMemcpyDeviceToHost, p->s_out);
    cudaMemcpyAsync(s.h_nm, s.d_nm, sizeof(int32_t) * (F - 1), cudaMemcpyDeviceToHost, p->s_out);
  }
  if (!p->check(cudaEventRecord(s.done, p->s_out), "event record") || !p->check(cudaGetLastError(), "pipeline submit"))
    return ORBX_E_CUDA;
  s.used = true;
  p->n_submitted = step + 1;
  return step;
}

int orbp_wait(orbp_pipeline* p, long long ticket, orbp_result* out) {
  if (!p || !out) return ORBX_E_INVALID;
  if (ticket < 0 || ticket >= p->n_submitted || ticket < p->n_submitted - p->cfg.depth) {
    p->err = "ticket no longer (or not yet) held by the pipeline";
    return ORBX_E_STATE;
  }
  Slot& s = p->slots[ticket % p->cfg.depth];
  if (!p->check(cudaEventSynchronize(s.done), "wait for step")) return ORBX_E_CUDA;
  for (int c = 0; c < ORBP_MAX_CAMS; ++c) {
    out->kps[c] = s.h_kps[c];
    out->desc[c] = s.h_desc[c];
    out->counts[c] = s.h_counts[c];
    out->cap[c] = p->caps[c];
  }
  out->matches12 = s.h_m12;
  out->nmatches = s.h_nm;
  out->rig_frames = p->cfg.rig_frames;
  return ORBX_OK;
}

int orbp_drain(orbp_pipeline* p) {
  if (!p) return ORBX_E_INVALID;
  for (cudaStream_t st : {p->s_in, p->s_compute, p->s_match, p->s_out})
    if (!p->check(cudaStreamSynchronize(st), "drain")) return ORBX_E_CUDA;
  return ORBX_OK;
}

void* orbp_stream(orbp_pipeline* p, int which) {
  if (!p) return nullptr;
  switch (which) {
    case 0: return (void*)p->s_in;
    case 1: return (void*)p->s_compute;
    case 2: return (void*)p->s_match;
    case 3: return (void*)p->s_out;
    default: return nullptr;
  }
}

long long orbp_launch_count(const orbp_pipeline* p) {
  if (!p) return 0;
  long long n = orbm_launch_count(p->matcher);
  for (int c = 0; c < p->cfg.n_cams; ++c) n += orbx_launch_count(p->ex[c]);
  return n;
}

#pragma GCC visibility pop
}
