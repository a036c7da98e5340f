// bb_gemv.cu -- memory-bound streaming kernels for decode-sized M (1..32).
//
// Replaces the reference's generated SIMT GEMV (bitblas/ops/general_matmul/tilelang/dequantize/
// gemv_dequantize_simt.py:164-262: 128-bit A loads, 4-byte packed-B loads, LOP3 decode, fp16 FMA /
// __dp4a, shuffle all-reduce).  B200 design:
//   * the packed weights are the only HBM stream that matters (N*K*bits/8 bytes); every warp owns 16
//     weight rows and a contiguous K range, reads them with 128-bit ld.global.nc.L1::no_allocate, 4
//     steps (4 KB / warp) in flight, every CTA resident at once so the memory system balances the tail;
//   * LOP3 decode is done in registers straight into mma.sync fragments (m16n8k16 f16/bf16, m16n8k32
//     u8.s8): 16 weight rows x 8 batch rows per instruction, fp32 / int32 accumulation, so the CUDA-core
//     pipes only see the decode (1 LOP3 + 1 HSUB2 per two weights), not the FMAs;
//   * k order inside a dot product is free, so the reference's interleaved storage layout AND the plain
//     compressed layout are both consumed with zero re-ordering cost (the B fragment is permuted instead);
//   * per-group scale / zero are applied to the group's partial sum (s * (sum(w*a) - z * sum(a))), the
//     integer part of z folded into the decode magic so GPTQ-style integer zero points cost nothing.
#include "bb_common.cuh"

namespace bb {

namespace {

constexpr int GEMV_WARPS = 4;
constexpr int PF = 4;  // weight steps in flight per warp

struct GemvParams {
  const void* A;
  const uint8_t* W;
  const void* scale;
  const void* zeros;
  const void* bias;
  void* C;
  int M, N, K;
  int g;          // group size in elements (multiple of 128, or K)
  int G;          // groups per row
  int with_scaling;
  int zmode;      // 0 none, 1 original, 2 rescale, 3 quantized
  int zp_const;   // constant zero point of the "int" formats (2^(bits-1)), 0 for uint
  int out_dtype;
  int rb_per_cta; // 16-row blocks per CTA
};

template <typename T>
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
template <>
__device__ __forceinline__ void mma_16816<__half>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <>
__device__ __forceinline__ void mma_16816<__nv_bfloat16>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_16832_u8s8(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <typename T>
__device__ __forceinline__ float ld_as_float(const void* p, size_t i) {
  return TypeTraits<T>::to_float(reinterpret_cast<const T*>(p)[i]);
}

template <int BITS>
struct WordsPerStep {  // 32 k per thread per step
  static constexpr int value = BITS;  // 4-bit: 4 words (16 B), 2-bit: 2 words (8 B)
};

template <int BITS>
__device__ __forceinline__ void load_w(const uint8_t* p, uint32_t (&w)[BITS]) {
  if constexpr (BITS == 4) {
    uint4 v = ldg_nc_v4(p);
    w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
  } else {
    uint2 v = ldg_nc_v2(p);
    w[0] = v.x; w[1] = v.y;
  }
}

// ---------------------------------------------------------------------------------------------
// fp16 / bf16 activations
// ---------------------------------------------------------------------------------------------
template <typename T, int BITS, bool IL, int NT>
__global__ void __launch_bounds__(GEMV_WARPS * 32)
gemv_mma_kernel(const GemvParams p) {
  constexpr int NP = 32 / BITS / 2;     // pairs per 32-bit word: 4 (4-bit) or 8 (2-bit)
  constexpr int WPS = BITS;             // words per thread per row per step
  constexpr uint32_t MAGIC = TypeTraits<T>::kMagic;
  constexpr int MAXZ = TypeTraits<T>::kMagicVal - (1 << BITS);  // largest zero point that folds exactly
  __shared__ float red[GEMV_WARPS][16][8 * NT];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = lane >> 2, q = lane & 3;
  const int nsteps_total = p.K / 128;
  const int spw = (nsteps_total + GEMV_WARPS - 1) / GEMV_WARPS;
  const int step_begin = warp * spw;
  const int step_end = min(nsteps_total, step_begin + spw);
  const int spg = p.g / 128;  // steps per group
  const size_t row_bytes = size_t(p.K) * BITS / 8;
  const T* Aptr = reinterpret_cast<const T*>(p.A);
  constexpr uint32_t ONE2 = std::is_same<T, __half>::value ? 0x3c003c00u : 0x3f803f80u;
  const uint32_t ones[4] = {ONE2, ONE2, ONE2, ONE2};

  for (int rbi = 0; rbi < p.rb_per_cta; ++rbi) {
    const int rb = blockIdx.x * p.rb_per_cta + rbi;
    if (rb * 16 >= p.N) break;
    const int n_a = rb * 16 + r, n_b = n_a + 8;
    const uint8_t* wrow_a = p.W + size_t(n_a) * row_bytes + q * (4 * WPS);
    const uint8_t* wrow_b = p.W + size_t(n_b) * row_bytes + q * (4 * WPS);

    float acc_t[NT][4], acc_g[NT][4], asum_g[NT][4];
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc_t[t][j] = acc_g[t][j] = asum_g[t][j] = 0.f;

    // per-group state
    float s_a = 1.f, s_b = 1.f, zc_a = 0.f, zc_b = 0.f;
    uint32_t mz_a = MAGIC + uint32_t(p.zp_const) * 0x00010001u, mz_b = mz_a;
    bool need_asum = false;

    auto begin_group = [&](int step) {
      const int gi = step / spg;
      if (p.with_scaling) {
        s_a = ld_as_float<T>(p.scale, size_t(n_a) * p.G + gi);
        s_b = ld_as_float<T>(p.scale, size_t(n_b) * p.G + gi);
      }
      if (p.zmode == 1 || p.zmode == 2) {
        const float za = ld_as_float<T>(p.zeros, size_t(n_a) * p.G + gi);
        const float zb = ld_as_float<T>(p.zeros, size_t(n_b) * p.G + gi);
        if (p.zmode == 1) {
          const float ia = rintf(za), ib = rintf(zb);
          const bool oka = ia >= 0.f && ia <= float(MAXZ), okb = ib >= 0.f && ib <= float(MAXZ);
          mz_a = MAGIC + (oka ? uint32_t(int(ia)) * 0x00010001u : 0u);
          mz_b = MAGIC + (okb ? uint32_t(int(ib)) * 0x00010001u : 0u);
          zc_a = (oka ? za - ia : za) * s_a;   // (w - z) * s = s*w - (s*z)
          zc_b = (okb ? zb - ib : zb) * s_b;
        } else {
          zc_a = za; zc_b = zb;                // w * s - z
        }
        need_asum = __any_sync(0xffffffffu, zc_a != 0.f || zc_b != 0.f);
      } else if (p.zmode == 3) {
        const uint8_t* qz = reinterpret_cast<const uint8_t*>(p.zeros) + size_t(gi) * (size_t(p.N) * BITS / 8);
        constexpr int EPB = 8 / BITS;
        const uint32_t za = (qz[n_a / EPB] >> (BITS * (n_a % EPB))) & ((1u << BITS) - 1u);
        const uint32_t zb = (qz[n_b / EPB] >> (BITS * (n_b % EPB))) & ((1u << BITS) - 1u);
        mz_a = MAGIC + za * 0x00010001u;
        mz_b = MAGIC + zb * 0x00010001u;
      }
    };
    auto end_group = [&]() {
#pragma unroll
      for (int t = 0; t < NT; ++t) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float s = (j < 2) ? s_a : s_b;
          float v = s * acc_g[t][j];
          if (need_asum) v -= ((j < 2) ? zc_a : zc_b) * asum_g[t][j & 1];
          acc_t[t][j] += v;
          acc_g[t][j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) asum_g[t][j] = 0.f;
      }
    };

    uint32_t wq[PF][2][WPS];
#pragma unroll
    for (int i = 0; i < PF; ++i) {
      if (step_begin + i < step_end) {
        load_w<BITS>(wrow_a + size_t(step_begin + i) * (16 * BITS), wq[i][0]);
        load_w<BITS>(wrow_b + size_t(step_begin + i) * (16 * BITS), wq[i][1]);
      }
    }

    bool group_open = false;
    for (int s0 = step_begin; s0 < step_end; s0 += PF) {
#pragma unroll
      for (int i = 0; i < PF; ++i) {
        const int step = s0 + i;
        if (step >= step_end) break;
        uint32_t wa[WPS], wb[WPS];
#pragma unroll
        for (int x = 0; x < WPS; ++x) { wa[x] = wq[i][0][x]; wb[x] = wq[i][1][x]; }
        if (step + PF < step_end) {
          load_w<BITS>(wrow_a + size_t(step + PF) * (16 * BITS), wq[i][0]);
          load_w<BITS>(wrow_b + size_t(step + PF) * (16 * BITS), wq[i][1]);
        }
        if (!group_open || step % spg == 0) {
          if (group_open) end_group();
          begin_group(step);
          group_open = true;
        }
        const int kq = step * 128 + q * 32;  // first k of this thread's 32-wide slice
#pragma unroll
        for (int wi = 0; wi < WPS; ++wi) {
          uint32_t ha[NP], hb[NP];
          if constexpr (BITS == 4) {
            decode_u4x8_raw<T>(wa[wi], ha);
            decode_u4x8_raw<T>(wb[wi], hb);
          } else if constexpr (IL) {
            decode_u2x16_raw_interleaved<T>(wa[wi], ha);
            decode_u2x16_raw_interleaved<T>(wb[wi], hb);
          } else {
            decode_u2x16_raw_compressed<T>(wa[wi], ha);
            decode_u2x16_raw_compressed<T>(wb[wi], hb);
          }
#pragma unroll
          for (int x = 0; x < NP; ++x) { ha[x] = sub2<T>(ha[x], mz_a); hb[x] = sub2<T>(hb[x], mz_b); }
          const int kw = kq + wi * (2 * NP);  // this word's first k
#pragma unroll
          for (int t = 0; t < NT; ++t) {
            uint32_t R[NP];
            const int m = 8 * t + r;
            if (m < p.M) {
              const uint4* ap = reinterpret_cast<const uint4*>(Aptr + size_t(m) * p.K + kw);
#pragma unroll
              for (int x = 0; x < NP / 4; ++x) {
                uint4 v = __ldg(ap + x);
                R[4 * x] = v.x; R[4 * x + 1] = v.y; R[4 * x + 2] = v.z; R[4 * x + 3] = v.w;
              }
            } else {
#pragma unroll
              for (int x = 0; x < NP; ++x) R[x] = 0u;
            }
#pragma unroll
            for (int j = 0; j < NP / 2; ++j) {
              const uint32_t af[4] = {ha[2 * j], hb[2 * j], ha[2 * j + 1], hb[2 * j + 1]};
              uint32_t b0, b1;
              if constexpr (IL) {
                b0 = R[2 * j]; b1 = R[2 * j + 1];
              } else {
                b0 = __byte_perm(R[j], R[j + NP / 2], 0x5410);
                b1 = __byte_perm(R[j], R[j + NP / 2], 0x7632);
              }
              mma_16816<T>(acc_g[t], af, b0, b1);
              if (need_asum) mma_16816<T>(asum_g[t], ones, b0, b1);
            }
          }
        }
      }
    }
    if (group_open) end_group();

    // cross-warp reduction + epilogue
    __syncthreads();  // protect `red` from the previous row block's readers
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      red[warp][r][8 * t + 2 * q] = acc_t[t][0];
      red[warp][r][8 * t + 2 * q + 1] = acc_t[t][1];
      red[warp][r + 8][8 * t + 2 * q] = acc_t[t][2];
      red[warp][r + 8][8 * t + 2 * q + 1] = acc_t[t][3];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 16 * 8 * NT; idx += GEMV_WARPS * 32) {
      const int row = idx & 15, m = idx >> 4;
      if (m >= p.M) continue;
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < GEMV_WARPS; ++w) v += red[w][row][m];
      const int n = rb * 16 + row;
      const size_t o = size_t(m) * p.N + n;
      if (p.out_dtype == BB_F16) {
        __half h = __float2half_rn(v);
        if (p.bias) h = __hadd(h, __float2half_rn(ld_as_float<T>(p.bias, n)));
        reinterpret_cast<__half*>(p.C)[o] = h;
      } else if (p.out_dtype == BB_BF16) {
        __nv_bfloat16 h = __float2bfloat16_rn(v);
        if (p.bias) h = __hadd(h, __float2bfloat16_rn(ld_as_float<T>(p.bias, n)));
        reinterpret_cast<__nv_bfloat16*>(p.C)[o] = h;
      } else {
        if (p.bias) v += ld_as_float<T>(p.bias, n);
        reinterpret_cast<float*>(p.C)[o] = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// int8 activations (W2A8 / W4A8), exact int32 accumulation.  Weights are used as raw unsigned fields
// (mma .u8.s8); the constant zero point of the signed formats is removed with zp * sum_k(a), the sum
// produced on the tensor cores with an all-ones A fragment.
// ---------------------------------------------------------------------------------------------
template <int BITS, int NT>
__global__ void __launch_bounds__(GEMV_WARPS * 32)
gemv_i8_kernel(const GemvParams p) {
  constexpr int WPS = BITS;            // words per thread per row per step (32 k)
  constexpr int RPW = 32 / BITS / 4;   // byte-quad registers per word: 4 (2-bit) or 2 (4-bit)
  __shared__ int red[GEMV_WARPS][16][8 * NT];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = lane >> 2, q = lane & 3;
  const int nsteps_total = p.K / 128;
  const int spw = (nsteps_total + GEMV_WARPS - 1) / GEMV_WARPS;
  const int step_begin = warp * spw;
  const int step_end = min(nsteps_total, step_begin + spw);
  const size_t row_bytes = size_t(p.K) * BITS / 8;
  const int8_t* Aptr = reinterpret_cast<const int8_t*>(p.A);
  const uint32_t ones[4] = {0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u};

  for (int rbi = 0; rbi < p.rb_per_cta; ++rbi) {
    const int rb = blockIdx.x * p.rb_per_cta + rbi;
    if (rb * 16 >= p.N) break;
    const int n_a = rb * 16 + r, n_b = n_a + 8;
    const uint8_t* wrow_a = p.W + size_t(n_a) * row_bytes + q * (4 * WPS);
    const uint8_t* wrow_b = p.W + size_t(n_b) * row_bytes + q * (4 * WPS);
    int acc[NT][4], asum[NT][4];
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[t][j] = asum[t][j] = 0;

    uint32_t wq[PF][2][WPS];
#pragma unroll
    for (int i = 0; i < PF; ++i) {
      if (step_begin + i < step_end) {
        load_w<BITS>(wrow_a + size_t(step_begin + i) * (16 * BITS), wq[i][0]);
        load_w<BITS>(wrow_b + size_t(step_begin + i) * (16 * BITS), wq[i][1]);
      }
    }
    for (int s0 = step_begin; s0 < step_end; s0 += PF) {
#pragma unroll
      for (int i = 0; i < PF; ++i) {
        const int step = s0 + i;
        if (step >= step_end) break;
        uint32_t da[8], db[8];  // 8 byte-quads = 32 k per row, natural k order
#pragma unroll
        for (int wi = 0; wi < WPS; ++wi) {
          if constexpr (BITS == 2) {
            uint32_t ta[4], tb[4];
            decode_u2x16_to_u8(wq[i][0][wi], 0u, ta);
            decode_u2x16_to_u8(wq[i][1][wi], 0u, tb);
#pragma unroll
            for (int x = 0; x < 4; ++x) { da[4 * wi + x] = ta[x]; db[4 * wi + x] = tb[x]; }
          } else {
            uint32_t ta[2], tb[2];
            decode_u4x8_to_u8(wq[i][0][wi], 0u, ta);
            decode_u4x8_to_u8(wq[i][1][wi], 0u, tb);
#pragma unroll
            for (int x = 0; x < 2; ++x) { da[2 * wi + x] = ta[x]; db[2 * wi + x] = tb[x]; }
          }
        }
        if (step + PF < step_end) {
          load_w<BITS>(wrow_a + size_t(step + PF) * (16 * BITS), wq[i][0]);
          load_w<BITS>(wrow_b + size_t(step + PF) * (16 * BITS), wq[i][1]);
        }
        const int kq = step * 128 + q * 32;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          uint32_t R[8];
          const int m = 8 * t + r;
          if (m < p.M) {
            const uint4* ap = reinterpret_cast<const uint4*>(Aptr + size_t(m) * p.K + kq);
            uint4 v0 = __ldg(ap), v1 = __ldg(ap + 1);
            R[0] = v0.x; R[1] = v0.y; R[2] = v0.z; R[3] = v0.w;
            R[4] = v1.x; R[5] = v1.y; R[6] = v1.z; R[7] = v1.w;
          } else {
#pragma unroll
            for (int x = 0; x < 8; ++x) R[x] = 0u;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t af[4] = {da[2 * j], db[2 * j], da[2 * j + 1], db[2 * j + 1]};
            mma_16832_u8s8(acc[t], af, R[2 * j], R[2 * j + 1]);
            if (p.zp_const) mma_16832_u8s8(asum[t], ones, R[2 * j], R[2 * j + 1]);
          }
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < NT; ++t) {
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[t][j] -= p.zp_const * asum[t][j & 1];
      red[warp][r][8 * t + 2 * q] = acc[t][0];
      red[warp][r][8 * t + 2 * q + 1] = acc[t][1];
      red[warp][r + 8][8 * t + 2 * q] = acc[t][2];
      red[warp][r + 8][8 * t + 2 * q + 1] = acc[t][3];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 16 * 8 * NT; idx += GEMV_WARPS * 32) {
      const int row = idx & 15, m = idx >> 4;
      if (m >= p.M) continue;
      int v = 0;
#pragma unroll
      for (int w = 0; w < GEMV_WARPS; ++w) v += red[w][row][m];
      const int n = rb * 16 + row;
      const int b = p.bias ? int(reinterpret_cast<const int8_t*>(p.bias)[n]) : 0;
      const size_t o = size_t(m) * p.N + n;
      switch (p.out_dtype) {
        case BB_I32: reinterpret_cast<int*>(p.C)[o] = v + b; break;
        case BB_I8: reinterpret_cast<int8_t*>(p.C)[o] = int8_t(int8_t(v) + b); break;
        case BB_F32: reinterpret_cast<float*>(p.C)[o] = float(v) + float(b); break;
        case BB_F16: reinterpret_cast<__half*>(p.C)[o] = __hadd(__int2half_rn(v), __int2half_rn(b)); break;
        default: reinterpret_cast<__nv_bfloat16*>(p.C)[o] = __hadd(__int2bfloat16_rn(v), __int2bfloat16_rn(b));
      }
    }
  }
}

GemvParams make_params(const MatmulArgs& a) {
  GemvParams p;
  const bb_matmul_desc& d = a.d;
  p.A = a.A; p.W = (const uint8_t*)a.W; p.scale = d.with_scaling ? a.scale : nullptr;
  p.zeros = d.with_zeros ? a.zeros : nullptr; p.bias = d.with_bias ? a.bias : nullptr; p.C = a.C;
  p.M = a.m; p.N = d.N; p.K = d.K;
  p.g = a.gsize(); p.G = a.groups();
  p.with_scaling = d.with_scaling;
  p.zmode = d.with_zeros ? (d.zeros_mode + 1) : 0;
  p.zp_const = (d.w_fmt == BB_W_INT) ? (1 << (d.w_bits - 1)) : 0;
  p.out_dtype = d.out_dtype;
  p.rb_per_cta = 1;
  return p;
}

template <typename K>
int pick_rb(K kernel, int n_blocks16) {
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, GEMV_WARPS * 32, 0) != cudaSuccess || occ < 1) occ = 4;
  const int cap = device_sm_count() * occ;
  return (n_blocks16 + cap - 1) / cap;
}

template <typename T, int BITS, bool IL>
int launch_mma_nt(const MatmulArgs& a, GemvParams p) {
  const int nb = p.N / 16;
  const int nt = (p.M + 7) / 8;
#define BB_GEMV_LAUNCH(NTV)                                                              \
  {                                                                                      \
    auto k = gemv_mma_kernel<T, BITS, IL, NTV>;                                          \
    p.rb_per_cta = pick_rb(k, nb);                                                       \
    k<<<(nb + p.rb_per_cta - 1) / p.rb_per_cta, GEMV_WARPS * 32, 0, a.stream>>>(p);      \
  }
  if (nt <= 1) BB_GEMV_LAUNCH(1)
  else if (nt == 2) BB_GEMV_LAUNCH(2)
  else BB_GEMV_LAUNCH(4)
#undef BB_GEMV_LAUNCH
  BB_LAUNCH_CHECK();
  return 0;
}

}  // namespace

bool gemv_mma_supported(const bb_matmul_desc& d, int m) {
  if (m < 1 || m > 32) return false;
  if (d.a_dtype != BB_F16 && d.a_dtype != BB_BF16) return false;
  if (d.w_fmt != BB_W_UINT && d.w_fmt != BB_W_INT) return false;
  if (d.w_bits != 4 && d.w_bits != 2) return false;
  if (d.w_layout == BB_LAYOUT_INTERLEAVED_8) return false;
  if (d.N % 16 || d.K % 128) return false;
  const int g = d.group_size <= 0 ? d.K : d.group_size;
  if (g % 128 || d.K % g) return false;
  if (d.with_zeros && !d.with_scaling) return false;
  if (d.w_fmt == BB_W_INT && d.with_zeros) return false;
  if (d.out_dtype != BB_F16 && d.out_dtype != BB_BF16 && d.out_dtype != BB_F32) return false;
  if (d.with_zeros && d.zeros_mode == BB_ZEROS_QUANTIZED && (d.N * d.w_bits) % 8) return false;
  return true;
}

int launch_gemv_mma(const MatmulArgs& a) {
  GemvParams p = make_params(a);
  const bool il = a.d.w_layout == BB_LAYOUT_INTERLEAVED_16;
  const bool f16 = a.d.a_dtype == BB_F16;
  const int bits = a.d.w_bits;
  if (f16) {
    if (bits == 4) return il ? launch_mma_nt<__half, 4, true>(a, p) : launch_mma_nt<__half, 4, false>(a, p);
    return il ? launch_mma_nt<__half, 2, true>(a, p) : launch_mma_nt<__half, 2, false>(a, p);
  }
  if (bits == 4) return il ? launch_mma_nt<__nv_bfloat16, 4, true>(a, p) : launch_mma_nt<__nv_bfloat16, 4, false>(a, p);
  return il ? launch_mma_nt<__nv_bfloat16, 2, true>(a, p) : launch_mma_nt<__nv_bfloat16, 2, false>(a, p);
}

bool gemv_i8_supported(const bb_matmul_desc& d, int m) {
  if (m < 1 || m > 32) return false;
  if (d.a_dtype != BB_I8 || d.accum_dtype != BB_I32) return false;
  if (d.w_fmt != BB_W_UINT && d.w_fmt != BB_W_INT) return false;
  if (d.w_bits != 4 && d.w_bits != 2) return false;
  if (d.w_layout != BB_LAYOUT_INTERLEAVED_8) return false;
  if (d.with_scaling || d.with_zeros) return false;
  if (d.N % 16 || d.K % 128) return false;
  return true;
}

int launch_gemv_i8(const MatmulArgs& a) {
  GemvParams p = make_params(a);
  const int nb = p.N / 16;
  const int nt = (p.M + 7) / 8;
#define BB_GEMV_I8_LAUNCH(BITSV, NTV)                                                    \
  {                                                                                      \
    auto k = gemv_i8_kernel<BITSV, NTV>;                                                 \
    p.rb_per_cta = pick_rb(k, nb);                                                       \
    k<<<(nb + p.rb_per_cta - 1) / p.rb_per_cta, GEMV_WARPS * 32, 0, a.stream>>>(p);      \
  }
  if (a.d.w_bits == 2) {
    if (nt <= 1) BB_GEMV_I8_LAUNCH(2, 1) else if (nt == 2) BB_GEMV_I8_LAUNCH(2, 2) else BB_GEMV_I8_LAUNCH(2, 4)
  } else {
    if (nt <= 1) BB_GEMV_I8_LAUNCH(4, 1) else if (nt == 2) BB_GEMV_I8_LAUNCH(4, 2) else BB_GEMV_I8_LAUNCH(4, 4)
  }
#undef BB_GEMV_I8_LAUNCH
  BB_LAUNCH_CHECK();
  return 0;
}

}  // namespace bb
