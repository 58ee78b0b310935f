// Dev tool: times the phases of ONE octree CTA (clock64 marks inside octree_core.h when
// OT_MARKS is defined) on candidates dumped by tools/dev/dump_candidates.py.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define OT_MARKS 1
__device__ long long* g_marks;
__device__ int g_nmarks;
#define OT_MARK(id) do { __syncthreads(); if (threadIdx.x == 0 && blockIdx.x == 0 && g_nmarks < 4096) { g_marks[2 * g_nmarks] = (id); g_marks[2 * g_nmarks + 1] = clock64(); ++g_nmarks; } } while (0)
#include "../../multi_orb_slam_b200/csrc/octree_core.h"

__global__ void __launch_bounds__(256) k(const uint32_t* keys_in, int M, OtRoots roots, int N, int cap, int scap, int smem_keys, uint32_t* out, int* nout, int threads) {
  OtScratch s;
  const int sb = ot_layout(s, cap, scap, 256);
  uint32_t* skeys = reinterpret_cast<uint32_t*>(ot_smem + sb);
  uint16_t* sknode = reinterpret_cast<uint16_t*>(ot_smem + sb + 4 * smem_keys);
  for (int i = threadIdx.x; i < M; i += blockDim.x) skeys[i] = keys_in[i];
  __syncthreads();
  OT_MARK(0);
  int n = ot_distribute(skeys, sknode, M, roots, N, s, out + blockIdx.x * cap);
  OT_MARK(99);
  if (threadIdx.x == 0) nout[blockIdx.x] = n;
}

int main(int argc, char** argv) {
  FILE* f = fopen(argv[1], "rb");
  int hdr[4];
  fread(hdr, 4, 4, f);
  int M = hdr[0], W = hdr[1], H = hdr[2], N = hdr[3];
  std::vector<uint32_t> keys(M);
  fread(keys.data(), 4, M, f);
  fclose(f);
  int nblocks = argc > 2 ? atoi(argv[2]) : 1;
  OtRoots roots;
  roots.n_ini = (int)std::round((float)W / (float)H);
  roots.hx = (float)W / roots.n_ini;
  for (int i = 0; i <= roots.n_ini; ++i) roots.root_x[i] = (int)(roots.hx * (float)i);
  roots.height = H;
  int cap = std::max(N + 3, 4 * roots.n_ini) + 1, scap = cap + 1, smem_keys = 6144;
  OtScratch s;
  size_t smem = ot_layout(s, cap, scap, 256) + (size_t)smem_keys * 6;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  uint32_t *dk, *dout; int* dn; long long* dm;
  cudaMalloc(&dk, 4 * M); cudaMalloc(&dout, 4 * cap * nblocks); cudaMalloc(&dn, 4 * nblocks); cudaMalloc(&dm, 16 * 4096);
  cudaMemcpy(dk, keys.data(), 4 * M, cudaMemcpyHostToDevice);
  cudaMemcpyToSymbol(g_marks, &dm, sizeof(dm));
  for (int rep = 0; rep < 3; ++rep) {
    int zero = 0;
    cudaMemcpyToSymbol(g_nmarks, &zero, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<<<nblocks, 256, smem>>>(dk, M, roots, N, cap, scap, smem_keys, dout, dn, 256);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int n; cudaMemcpy(&n, dn, 4, cudaMemcpyDeviceToHost);
    printf("rep %d: %s, M=%d N=%d -> %d keypoints, kernel %.1f us (%d blocks)\n", rep, cudaGetErrorString(e), M, N, n, ms * 1e3, nblocks);
  }
  int nm; cudaMemcpyFromSymbol(&nm, g_nmarks, 4);
  std::vector<long long> marks(2 * nm);
  cudaMemcpy(marks.data(), dm, 16 * nm, cudaMemcpyDeviceToHost);
  // aggregate cycles by (from id -> to id)
  long long agg[128] = {0}; int cnt[128] = {0};
  for (int i = 1; i < nm; ++i) { int id = (int)marks[2 * i]; agg[id] += marks[2 * i + 1] - marks[2 * i - 1]; cnt[id]++; }
  printf("marks: %d, total cycles %lld\n", nm, marks[2 * nm - 1] - marks[1]);
  for (int id = 0; id < 128; ++id) if (cnt[id]) printf("  phase ending at mark %3d: %3d times, %8lld cycles total, %7lld avg\n", id, cnt[id], agg[id], agg[id] / cnt[id]);
  return 0;
}
