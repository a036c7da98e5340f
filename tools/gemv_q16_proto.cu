// tools/gemv_q16_proto.cu -- round-2 PROTOTYPE (never run yet; self-checking): W4A16 m = 1 GEMV on INTEGER tensor cores.
// The fp16 activations are quantised once per call to 16-bit fixed point (per-token scale amax / 32767) and split into a signed high
// byte and an unsigned low byte; the u4 weights become u8 with ONE LOP3 per four weights (w & 0x0f0f0f0f, (w >> 4) & 0x0f0f0f0f)
// and go through mma.sync.m16n8k32 (u8 x s8 for the high plane, u8 x u8 for the low plane), exact int32 accumulation.  The
// zero point needs no second MMA: sum_k (u - z) a = sum_k u a - z * sum_k a, with the per-group sums of the quantised activations
// produced by the pre-pass.  The activation planes are stored in the weights' nibble order (interleaved-16 layout: byte order
// e0,e4,e1,e5 | e2,e6,e3,e7 inside every 8-element word), so no permute is needed in the hot loop.
// Same launch structure as gemv_mma_kernel (CTA = 16 rows, 4 warps split K, 4-step register queue); uint4 g = 128, quantized zeros.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/gemv_q16_proto tools/gemv_q16_proto.cu
// Run:   tools/gemv_q16_proto [N] [K]        (N % 16 == 0, K % 512 == 0)
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

// ---- pre-pass: one CTA.  a_hi / a_lo: [K] bytes in nibble order, sumq: [K/128] int32, sa: 1 float ----
__global__ void __launch_bounds__(1024) act_quant_kernel(const __half* A, int K, int8_t* a_hi, uint8_t* a_lo, int* sumq, float* sa) {
  __shared__ float smax[32];
  __shared__ float s_scale;
  float m = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) m = fmaxf(m, fabsf(__half2float(A[k])));
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = smax[threadIdx.x];
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (threadIdx.x == 0) { s_scale = v > 0.f ? v / 32767.f : 1.f; *sa = s_scale; }
  }
  __syncthreads();
  const float inv = 1.f / s_scale;
  // one thread per 8-element word; 16 consecutive threads cover one group of 128
  for (int wd = threadIdx.x; wd < K / 8; wd += blockDim.x) {
    int q[8], s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { q[i] = __float2int_rn(__half2float(A[wd * 8 + i]) * inv); s += q[i]; }
    const int order[8] = {0, 4, 1, 5, 2, 6, 3, 7};   // byte j of the word's 8-byte group holds element order[j]
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a_hi[wd * 8 + j] = int8_t(q[order[j]] >> 8);
      a_lo[wd * 8 + j] = uint8_t(q[order[j]] & 255);
    }
    for (int o = 8; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);   // 16 lanes = one group (K/8 is a multiple of 16)
    if ((threadIdx.x & 15) == 0) sumq[wd / 16] = s;
  }
}

__device__ __forceinline__ uint4 ldg_nc(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void imma_u8s8(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void imma_u8u8(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int KS = 4;   // warps per CTA (K splits)
constexpr int PF = 4;   // steps in flight

__global__ void __launch_bounds__(KS * 32)
gemv_q16_kernel(const uint8_t* __restrict__ W, const __half* __restrict__ S, const uint8_t* __restrict__ Z, const int8_t* __restrict__ a_hi,
                const uint8_t* __restrict__ a_lo, const int* __restrict__ sumq, const float* __restrict__ sa, __half* __restrict__ C, int N, int K) {
  __shared__ float red[KS][16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, r = lane >> 2, q = lane & 3;
  const int rb = blockIdx.x, G = K / 128;
  const int steps = G, per = steps / KS, s0 = warp * per;   // host guarantees G % KS == 0 and per % PF == 0
  const size_t row_bytes = size_t(K) / 2;
  const int n_a = rb * 16 + r, n_b = n_a + 8;
  const uint8_t* wa = W + size_t(n_a) * row_bytes + size_t(s0) * 64 + q * 16;
  const uint8_t* wb = wa + 8 * row_bytes;
  const int8_t* ah = a_hi + size_t(s0) * 128 + q * 32;
  const uint8_t* al = a_lo + size_t(s0) * 128 + q * 32;
  const __half* sp_a = S + size_t(n_a) * G + s0;
  const __half* sp_b = S + size_t(n_b) * G + s0;
  const uint8_t* zp = Z + size_t(s0) * (N / 2) + n_a / 2;   // rows n_a and n_b = n_a + 8: bytes +0 and +4, same nibble
  const uint32_t zsh = 4u * uint32_t(n_a & 1);
  const int* sq = sumq + s0;
  float acc_t[4] = {0.f, 0.f, 0.f, 0.f};
  uint4 qa[PF], qb[PF];
#pragma unroll
  for (int u = 0; u < PF; ++u) { qa[u] = ldg_nc(wa + u * 64); qb[u] = ldg_nc(wb + u * 64); }
  for (int s = 0; s < per; s += PF) {
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const uint32_t xa[4] = {qa[u].x, qa[u].y, qa[u].z, qa[u].w}, xb[4] = {qb[u].x, qb[u].y, qb[u].z, qb[u].w};
      if (s + u + PF < per) { qa[u] = ldg_nc(wa + (s + u + PF) * 64); qb[u] = ldg_nc(wb + (s + u + PF) * 64); }
      const uint4 h0 = __ldg(reinterpret_cast<const uint4*>(ah + (s + u) * 128)), h1 = __ldg(reinterpret_cast<const uint4*>(ah + (s + u) * 128) + 1);
      const uint4 l0 = __ldg(reinterpret_cast<const uint4*>(al + (s + u) * 128)), l1 = __ldg(reinterpret_cast<const uint4*>(al + (s + u) * 128) + 1);
      const uint32_t Rh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w}, Rl[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
      const float s_a = __half2float(__ldg(sp_a + s + u)), s_b = __half2float(__ldg(sp_b + s + u));
      const uint32_t za = (uint32_t(__ldg(zp + size_t(s + u) * (N / 2))) >> zsh) & 15u;
      const uint32_t zb = (uint32_t(__ldg(zp + size_t(s + u) * (N / 2) + 4)) >> zsh) & 15u;
      const int sum_g = __ldg(sq + s + u);
      int acc_h[4] = {0, 0, 0, 0}, acc_l[4] = {0, 0, 0, 0};
#pragma unroll
      for (int wi = 0; wi < 4; ++wi) {
        const uint32_t af[4] = {xa[wi] & 0x0f0f0f0fu, xb[wi] & 0x0f0f0f0fu, (xa[wi] >> 4) & 0x0f0f0f0fu, (xb[wi] >> 4) & 0x0f0f0f0fu};
        imma_u8s8(acc_h, af, Rh[2 * wi], Rh[2 * wi + 1]);
        imma_u8u8(acc_l, af, Rl[2 * wi], Rl[2 * wi + 1]);
      }
      // batch column 0 only (m = 1): accumulators 0 (row n_a) and 2 (row n_b) of the q == 0 lanes; computed by all lanes
      acc_t[0] = fmaf(s_a, float(acc_h[0] * 256 + acc_l[0] - int(za) * sum_g), acc_t[0]);
      acc_t[2] = fmaf(s_b, float(acc_h[2] * 256 + acc_l[2] - int(zb) * sum_g), acc_t[2]);
    }
  }
  if (q == 0) { red[warp][r] = acc_t[0]; red[warp][r + 8] = acc_t[2]; }
  __syncthreads();
  if (threadIdx.x < 16) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < KS; ++k) v += red[k][threadIdx.x];
    C[rb * 16 + threadIdx.x] = __float2half_rn(v * __ldg(sa));
  }
}

// naive fp32 reference on the ORIGINAL fp16 activations (interleaved-16 nibble order: nibble j = element 2j, nibble j+4 = element 2j+1)
__global__ void ref_kernel(const uint32_t* W, const __half* S, const uint8_t* Z, const __half* A, float* out, int N, int K) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int G = K / 128;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) {
    const uint32_t word = W[size_t(n) * (K / 8) + k / 8];
    const int e = k % 8, nib = (e & 1) ? (e / 2 + 4) : (e / 2);
    const int u = (word >> (4 * nib)) & 15, g = k / 128;
    const int zq = (Z[size_t(g) * (N / 2) + n / 2] >> (4 * (n & 1))) & 15;
    acc += float(u - zq) * __half2float(S[size_t(n) * G + g]) * __half2float(A[k]);
  }
  out[n] = acc;
}

int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 12288, K = argc > 2 ? atoi(argv[2]) : 12288;
  const int G = K / 128;
  if (N % 16 || K % 128 || G % KS || (G / KS) % PF) { printf("need N %% 16 == 0 and K / 128 divisible by %d\n", KS * PF); return 1; }
  const size_t wbytes = size_t(N) * K / 2;
  const int NCOPY = 5;
  std::vector<uint8_t*> Ws(NCOPY);
  std::vector<uint8_t> hw(wbytes);
  srand(1);
  for (auto& b : hw) b = uint8_t(rand());
  for (auto& p : Ws) { CK(cudaMalloc(&p, wbytes)); CK(cudaMemcpy(p, hw.data(), wbytes, cudaMemcpyHostToDevice)); }
  std::vector<__half> hs(size_t(N) * G), ha(K);
  for (auto& v : hs) v = __float2half(0.002f + 0.02f * (rand() % 1000) / 1000.f);
  for (auto& v : ha) v = __float2half((rand() % 2000) / 2000.f - 0.5f);
  ha[7] = __float2half(11.f);   // an outlier, as in LLM activations: the fixed-point step is amax / 32767
  std::vector<uint8_t> hz(size_t(G) * N / 2);
  for (auto& b : hz) b = uint8_t(rand());
  __half *dS, *dA, *dC; uint8_t *dZ, *dLo; int8_t* dHi; int* dSum; float *dSa, *dRef;
  CK(cudaMalloc(&dS, hs.size() * 2)); CK(cudaMemcpy(dS, hs.data(), hs.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&dA, K * 2)); CK(cudaMemcpy(dA, ha.data(), K * 2, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&dZ, hz.size())); CK(cudaMemcpy(dZ, hz.data(), hz.size(), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&dC, N * 2)); CK(cudaMalloc(&dRef, N * 4)); CK(cudaMalloc(&dHi, K)); CK(cudaMalloc(&dLo, K)); CK(cudaMalloc(&dSum, G * 4)); CK(cudaMalloc(&dSa, 4));
  auto launch = [&](uint8_t* w) {
    act_quant_kernel<<<1, 1024>>>(dA, K, dHi, dLo, dSum, dSa);
    gemv_q16_kernel<<<N / 16, KS * 32>>>(w, dS, dZ, dHi, dLo, dSum, dSa, dC, N, K);
  };
  launch(Ws[0]);
  ref_kernel<<<(N + 127) / 128, 128>>>((const uint32_t*)Ws[0], dS, dZ, dA, dRef, N, K);
  CK(cudaDeviceSynchronize());
  std::vector<__half> hc(N); std::vector<float> hr(N);
  CK(cudaMemcpy(hc.data(), dC, N * 2, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hr.data(), dRef, N * 4, cudaMemcpyDeviceToHost));
  double num = 0, den = 0;
  for (int i = 0; i < N; ++i) { const double d = double(__half2float(hc[i])) - hr[i]; num += d * d; den += double(hr[i]) * hr[i]; }
  printf("N=%d K=%d  normwise rel err %.3e (%s)\n", N, K, std::sqrt(num / den), std::sqrt(num / den) < 1e-2 ? "ok" : "MISMATCH");
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 5; ++i) launch(Ws[i % NCOPY]);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(a);
  for (int i = 0; i < 40; ++i) launch(Ws[i % NCOPY]);
  cudaEventRecord(b); CK(cudaEventSynchronize(b));
  float ms; cudaEventElapsedTime(&ms, a, b); ms /= 40;
  const double bytes = double(wbytes) + double(N) * G * 2.5 + K * 2 + N * 2;
  printf("%.2f us per call (pre-pass + GEMV), %.0f GB/s\n", ms * 1e3, bytes / (ms * 1e-3) / 1e9);
  cudaEventRecord(a);
  for (int i = 0; i < 40; ++i) gemv_q16_kernel<<<N / 16, KS * 32>>>(Ws[i % NCOPY], dS, dZ, dHi, dLo, dSum, dSa, dC, N, K);
  cudaEventRecord(b); CK(cudaEventSynchronize(b));
  cudaEventElapsedTime(&ms, a, b); ms /= 40;
  printf("%.2f us per call (GEMV only), %.0f GB/s\n", ms * 1e3, bytes / (ms * 1e-3) / 1e9);
  return 0;
}
