// Does programmatic dependent launch overlap consecutive launches on this box —
// from a plain stream, and from a stream-captured CUDA graph?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o pdl_probe pdl_probe.cu && ./pdl_probe
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t;
}

// Each CTA: stamp entry, (optionally) griddepcontrol.wait, stamp, spin ~busy_ns, stamp exit.
__global__ void probe(unsigned long long* stamps, int k, int busy_ns, int use_pdl) {
  if (use_pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const unsigned long long t0 = gtime();
  if (use_pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
  const unsigned long long t1 = gtime();
  while (gtime() - t1 < (unsigned long long)busy_ns) { }
  const unsigned long long t2 = gtime();
  if (threadIdx.x == 0) {
    unsigned long long* s = stamps + ((size_t)k * gridDim.x + blockIdx.x) * 3;
    s[0] = t0; s[1] = t1; s[2] = t2;
  }
}

static void launch(cudaStream_t s, unsigned long long* d, int k, int grid, int busy, int pdl) {
  cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  if (pdl) { cfg.attrs = attr; cfg.numAttrs = 1; }
  cudaLaunchKernelEx(&cfg, probe, d, k, busy, pdl);
}

static void report(const char* what, unsigned long long* h, int K, int grid) {
  // per launch: first entry, first past-wait, last exit; gap = next.first_past_wait - this.last_exit
  double gap = 0, early = 0;
  for (int k = 0; k + 1 < K; ++k) {
    unsigned long long last_exit = 0, next_entry = ~0ull, next_wait = ~0ull;
    for (int b = 0; b < grid; ++b) {
      unsigned long long* s = h + ((size_t)k * grid + b) * 3;
      unsigned long long* n = h + ((size_t)(k + 1) * grid + b) * 3;
      if (s[2] > last_exit) last_exit = s[2];
      if (n[0] < next_entry) next_entry = n[0];
      if (n[1] < next_wait) next_wait = n[1];
    }
    gap += (double)((long long)next_wait - (long long)last_exit);
    early += (double)((long long)last_exit - (long long)next_entry);
  }
  printf("%-34s next kernel resumes %.0f ns after the previous one's last CTA exit; its first CTA was "
         "resident %.0f ns BEFORE that exit\n", what, gap / (K - 1), early / (K - 1));
}

int main() {
  const int K = 12, grid = 148, busy = 10000;
  unsigned long long *d, *h;
  const size_t bytes = (size_t)K * grid * 3 * sizeof(unsigned long long);
  cudaMalloc(&d, bytes); h = (unsigned long long*)malloc(bytes);
  cudaStream_t s; cudaStreamCreate(&s);
  for (int pdl = 0; pdl < 2; ++pdl) {
    for (int rep = 0; rep < 2; ++rep) {            // second pass is warm
      for (int k = 0; k < K; ++k) launch(s, d, k, grid, busy, pdl);
      cudaStreamSynchronize(s);
    }
    cudaMemcpy(h, d, bytes, cudaMemcpyDeviceToHost);
    report(pdl ? "stream, PDL attribute" : "stream, plain", h, K, grid);
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
    for (int k = 0; k < K; ++k) launch(s, d, k, grid, busy, pdl);
    cudaStreamEndCapture(s, &g);
    cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
    if (e != cudaSuccess) { printf("instantiate: %s\n", cudaGetErrorString(e)); continue; }
    for (int rep = 0; rep < 3; ++rep) { cudaGraphLaunch(ge, s); cudaStreamSynchronize(s); }
    cudaMemcpy(h, d, bytes, cudaMemcpyDeviceToHost);
    report(pdl ? "captured graph, PDL attribute" : "captured graph, plain", h, K, grid);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
