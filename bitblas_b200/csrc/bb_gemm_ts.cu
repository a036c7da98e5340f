// bb_gemm_ts.cu -- the tensor-core path: tcgen05.mma with the dequantised weights as the TMEM ("TS") operand.
//
// Replaces the reference's generated mma.sync GEMM (bitblas/ops/general_matmul/tilelang/dequantize/
// matmul_dequantize_mma.py:333-506: cp.async -> smem(int8) -> regs -> LOP3 decode -> smem(fp16) -> ldmatrix
// -> mma.sync m16n8k16).  On B200 the decoded weights never touch shared memory:
//
//     C^T[n, m] = sum_k  Wdeq[n, k] * A[m, k]            (the operands are swapped: W is the MMA "A"/M side)
//
//   TMA warp      : per k-block, one 2-D tile of activations A[BM x 64] (fp16/bf16; 128 int8) into 128B-swizzled
//                   shared memory (the MMA "B"/N operand, K-major) and one tile of PACKED weights
//                   W[128 rows x 32 B] (4-bit) -- 8x fewer bytes than a dequantised tile;
//   dequant warps : 8 warps, thread <-> weight row (= TMEM lane).  ld.shared 16 B -> LOP3 decode + (w-z)*s
//                   in registers -> tcgen05.st into a ring of TMEM operand slots (32 columns per k-block);
//   MMA warp      : one elected thread issues tcgen05.mma.kind::f16 (or kind::i8) M=128, N=BM, K=16 (32),
//                   A operand from TMEM, B operand from the swizzled smem descriptor, fp32 (s32) accumulator
//                   in TMEM; tcgen05.commit releases the smem stage and the TMEM slot through mbarriers;
//   epilogue      : the dequant warps read the accumulator with tcgen05.ld, cast, add bias, store C[m, n].
//
// Shared-memory traffic per k-block is therefore A-tile (TMA write + MMA read) + 4 KB packed W, instead of
// also writing and re-reading a 16-32 KB dequantised W tile -- the difference between fitting and not fitting
// the 128 B/clk/SM shared-memory budget at full tensor rate.
#include <cuda.h>

#include <algorithm>
#include <mutex>

#include "bb_common.cuh"

namespace bb {

namespace {

constexpr int TS_ROWS = 128;    // weight rows per CTA = MMA M
// 16 dequant warps = 4 TMEM lane quadrants x 4.  Two organisations (dq_groups(BM)):
//  * BM > 128 (one CTA per SM, 36 KB stages, 6 of them): 2 groups x (4 quadrants x 2 column halves); a thread decodes HALF a row
//    of its group's k-blocks.  A k-block leaves the dequant stage quickly, so at most two smem stages are pinned by decoding.
//  * BM <= 128 (small stages, 8 of them): 4 groups x 4 quadrants; a thread decodes its WHOLE row (both halves) per iteration,
//    so the barrier / slot bookkeeping -- about as many instructions as one half's decode -- is paid once per 64 weights
//    (m = 16: 47 -> 38 us; with the large tile this organisation pins 4 of 6 stages and starves the TMA ring: 1285 -> 866 TFLOPS).
constexpr int DQ_WARPS = 16;
__host__ __device__ constexpr int dq_groups(int BM) { return BM > 128 ? 2 : 4; }
constexpr int TS_THREADS = (2 + DQ_WARPS) * 32;
constexpr int TA_SLOTS = 4;     // TMEM operand slots (2 per dequant group with 2 groups, 1 per group with 4)

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem desc]
template <bool INT8>
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (INT8) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}

// K-major, 128B-swizzled shared-memory operand descriptor (rows of 128 B, 8-row swizzle atoms = 1024 B)
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr & 0x3FFFFu) >> 4);  // start address, 16-byte units
  d |= uint64_t(0) << 16;                      // leading byte offset: unused for swizzled K-major
  d |= uint64_t(1024 >> 4) << 32;              // stride byte offset: 8 rows * 128 B
  d |= uint64_t(1) << 46;                      // descriptor version (sm_100)
  d |= uint64_t(2) << 61;                      // SWIZZLE_128B
  return d;
}

// ---------------------------------------------------------------------------------------------
struct TsParams {
  const void* scale;
  const void* zeros;
  const void* bias;
  OutSpec out;
  int M, N, K;
  int g, G;
  int mode;       // 0 none, 1 scale, 2 original, 3 rescale, 4 quantized
  int zp_const;
  int out_dtype;
  int a_dtype;
  int m_tiles;
  int stages;        // smem pipeline depth (even; <= TsSmem::kStages)
  int splits;        // split-K factor (>1: fp32 / int32 partials go to `ws`, reduced by splitk_reduce_kernel)
  int kb_per_split;  // k-blocks per split
  void* ws;          // [splits][M][N] partials
  const void* lut;   // NF4 table (16 x A_dtype), null otherwise
  int w_fmt;         // bb_wfmt (the table formats run the FMT = 1 instances)
  int w_tiled;       // BB_TILE_SLAB weight storage: tmW is the 4-D map {512 B, 32 rows, segments per row, row blocks}
  int staged_epi;    // 16-bit outputs: transpose the accumulator tile through shared memory, 16-byte row-major stores (see epilogue)
};

template <typename T>
struct ElemInfo {
  static constexpr bool kInt8 = false;
  static constexpr int kKB = 64;  // elements per 128-byte swizzle row
};
template <>
struct ElemInfo<int8_t> {
  static constexpr bool kInt8 = true;
  static constexpr int kKB = 128;
};

template <int BM>
struct TsTmem {
  static constexpr int kAcol0 = BM < 32 ? 32 : BM;
  static constexpr int kSlots = TA_SLOTS;
  static constexpr int kNeed = kAcol0 + kSlots * 32;
  static constexpr int kCols = kNeed <= 64 ? 64 : (kNeed <= 128 ? 128 : (kNeed <= 256 ? 256 : 512));
};

template <typename T, int BITS, int BM>
struct TsSmem {
  static constexpr int kPRB = ElemInfo<T>::kKB * BITS / 8;  // packed bytes per row per k-block
  static constexpr int kActBytes = BM * 128;
  static constexpr int kWBytes = TS_ROWS * kPRB;
  static constexpr int kStageBytes = kActBytes + kWBytes;
  static constexpr int kStagesRaw = (220 * 1024) / kStageBytes;
  static constexpr int kStages = (kStagesRaw > 8 ? 8 : kStagesRaw) & ~1;  // even: stage parity is static per dequant group
  static constexpr int kBarBytes = 512;   // full[8] + empty[8] + a_ready[8] + a_free[8] + acc_full + tmem slot word
  static constexpr int kTotal = kStages * kStageBytes + kBarBytes + 1024;  // + alignment slack
};

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}

// decode this thread's half of the k-block (PRB/2 packed bytes at shared address `src`) into 16 registers of
// natural-k-order 16-bit operand pairs.  MODE: 0 none, 1 scale, 2 original, 3 rescale, 4 quantized.
// IL = LOP3-interleaved storage (fast_decoding); !IL = plain compressed storage, re-ordered with byte permutes.
template <typename T, int BITS, int MODE, bool IL>
__device__ __forceinline__ void dequant_half_row(uint32_t src, uint32_t (&out)[16], const DqConst& c) {
  constexpr bool HI = IL && std::is_same<T, __half>::value && BITS == 4;
  constexpr uint32_t M = TypeTraits<T>::kMagic;
  auto fin = [&](uint32_t x, uint32_t mz) -> uint32_t { return dq_finish<T, MODE, BITS>(x, mz, c); };
  if constexpr (BITS == 4) {
    const uint4 pk = lds128(src);
    const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if constexpr (HI) {
        // odd nibbles extracted in place: mantissa bits 4..7 under exponent 2^6 (0x5400) read exactly 64 + u
        const uint32_t x = w[i], y = w[i] >> 8;
        out[4 * i + 0] = fin(lop3_and_or(x, 0x000f000fu, M), c.mz_lo);
        out[4 * i + 1] = fin(lop3_and_or(x, 0x00f000f0u, 0x54005400u), c.mz_hi);
        out[4 * i + 2] = fin(lop3_and_or(y, 0x000f000fu, M), c.mz_lo);
        out[4 * i + 3] = fin(lop3_and_or(y, 0x00f000f0u, 0x54005400u), c.mz_hi);
      } else if constexpr (IL) {
        uint32_t h[4];
        decode_u4x8_raw<T>(w[i], h);
#pragma unroll
        for (int j = 0; j < 4; ++j) out[4 * i + j] = fin(h[j], c.mz_lo);
      } else {
        // compressed: nibbles e0..e7 in order.  bytes t = (e0,e2,e4,e6), u = (e1,e3,e5,e7)
        constexpr uint32_t MB = (M >> 8) & 0xffu, MBYTES = MB * 0x01010101u;  // magic high byte (0x64 / 0x43)
        const uint32_t t = w[i] & 0x0f0f0f0fu, u = (w[i] >> 4) & 0x0f0f0f0fu;
        const uint32_t x = __byte_perm(t, u, 0x5140), y = __byte_perm(t, u, 0x7362);  // (e0,e1,e2,e3), (e4,e5,e6,e7)
        out[4 * i + 0] = fin(__byte_perm(x, MBYTES, 0x4140), c.mz_lo);
        out[4 * i + 1] = fin(__byte_perm(x, MBYTES, 0x4342), c.mz_lo);
        out[4 * i + 2] = fin(__byte_perm(y, MBYTES, 0x4140), c.mz_lo);
        out[4 * i + 3] = fin(__byte_perm(y, MBYTES, 0x4342), c.mz_lo);
      }
    }
  } else {
    const uint2 pk = lds64(src);
    const uint32_t w[2] = {pk.x, pk.y};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if constexpr (IL) {
        uint32_t h[8];
        decode_u2x16_raw_interleaved<T>(w[i], h);
#pragma unroll
        for (int j = 0; j < 8; ++j) out[8 * i + j] = fin(h[j], c.mz_lo);
      } else {
        // compressed: fields e0..e15 in order.  byte planes s_k = (e_k, e_{k+4}, e_{k+8}, e_{k+12})
        constexpr uint32_t MB = (M >> 8) & 0xffu, MBYTES = MB * 0x01010101u;
        const uint32_t s0 = w[i] & 0x03030303u, s1 = (w[i] >> 2) & 0x03030303u, s2 = (w[i] >> 4) & 0x03030303u,
                       s3 = (w[i] >> 6) & 0x03030303u;
        const uint32_t a01 = __byte_perm(s0, s1, 0x5140), a23 = __byte_perm(s2, s3, 0x5140);  // (e0,e1,e4,e5) (e2,e3,e6,e7)
        const uint32_t b01 = __byte_perm(s0, s1, 0x7362), b23 = __byte_perm(s2, s3, 0x7362);  // (e8,e9,e12,e13) ...
        out[8 * i + 0] = fin(__byte_perm(a01, MBYTES, 0x4140), c.mz_lo);
        out[8 * i + 1] = fin(__byte_perm(a23, MBYTES, 0x4140), c.mz_lo);
        out[8 * i + 2] = fin(__byte_perm(a01, MBYTES, 0x4342), c.mz_lo);
        out[8 * i + 3] = fin(__byte_perm(a23, MBYTES, 0x4342), c.mz_lo);
        out[8 * i + 4] = fin(__byte_perm(b01, MBYTES, 0x4140), c.mz_lo);
        out[8 * i + 5] = fin(__byte_perm(b23, MBYTES, 0x4140), c.mz_lo);
        out[8 * i + 6] = fin(__byte_perm(b01, MBYTES, 0x4342), c.mz_lo);
        out[8 * i + 7] = fin(__byte_perm(b23, MBYTES, 0x4342), c.mz_lo);
      }
    }
  }
}

// table formats (NF4 / fp4), compressed storage: 16 packed bytes -> 32 values in natural k order; MODE 0 none, 1 scale
template <typename T, int MODE>
__device__ __forceinline__ void dequant_half_row_lut(uint32_t src, uint32_t (&out)[16], const Lut4& t, uint32_t s2) {
  const uint4 pk = lds128(src);
  const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t h[4];
    lut4_decode8(w[i], t, h);
#pragma unroll
    for (int j = 0; j < 4; ++j) out[4 * i + j] = MODE == 1 ? mul2<T>(h[j], s2) : h[j];
  }
}

// 8-bit float weights (e4m3_float8 / e5m2_float8 with 16-bit activations): 32 bytes -> 32 values in natural k order.
// e5m2 is the top byte of an fp16 (quantization.py:179-182); e4m3 follows the reference's bit trick (quantization.py:169-176:
// sign | ((v & 63) << 7 | e4 << 8 | e4 << 7) ^ 0x2000 with e4 = v & 0x40), evaluated on two bytes at a time.  bf16 activations:
// the fp16 pair is widened exactly (every fp8 value is representable in bf16).
template <typename T, int MODE>
__device__ __forceinline__ void dequant_half_row_fp8(uint32_t src, uint32_t (&out)[16], bool e5m2, uint32_t s2) {
  const uint4 p0 = lds128(src), p1 = lds128(src + 16);
  const uint32_t w[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const uint32_t x = __byte_perm(w[i], 0u, j ? 0x3424 : 0x1404);   // (byte 2j) << 8 | (byte 2j+1) << 24
      uint32_t h = x;
      if (!e5m2) h = ((x & 0xC000C000u) | ((x >> 1) & 0x3F803F80u)) ^ 0x20002000u;
      if constexpr (std::is_same<T, __nv_bfloat16>::value) {
        const float2 f = __half22float2(u32_as_h2(h));
        h = b2_as_u32(__floats2bfloat162_rn(f.x, f.y));
      }
      out[2 * i + j] = MODE == 1 ? mul2<T>(h, s2) : h;
    }
  }
}

template <int BITS, bool IL>
__device__ __forceinline__ void dequant_half_row_i8(uint32_t src, uint32_t (&out)[16], uint32_t zp4) {
  if constexpr (BITS == 2) {
    const uint4 pk = lds128(src);
    const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint32_t h[4];
      if constexpr (IL) {
        decode_u2x16_to_u8(w[i], zp4 ? 0x80808080u : 0u, h);
      } else {
        const uint32_t orv = zp4 ? 0x80808080u : 0u;
        const uint32_t s0 = lop3_and_or(w[i], 0x03030303u, orv), s1 = lop3_and_or(w[i] >> 2, 0x03030303u, orv),
                       s2 = lop3_and_or(w[i] >> 4, 0x03030303u, orv), s3 = lop3_and_or(w[i] >> 6, 0x03030303u, orv);
        const uint32_t a = __byte_perm(s0, s1, 0x5140), b = __byte_perm(s2, s3, 0x5140);
        const uint32_t a2 = __byte_perm(s0, s1, 0x7362), b2 = __byte_perm(s2, s3, 0x7362);
        h[0] = __byte_perm(a, b, 0x5410); h[1] = __byte_perm(a, b, 0x7632);
        h[2] = __byte_perm(a2, b2, 0x5410); h[3] = __byte_perm(a2, b2, 0x7632);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) out[4 * i + j] = zp4 ? bytes_sub_zp(h[j], zp4) : h[j];
    }
  } else {
    const uint4 p0 = lds128(src);
    const uint4 p1 = lds128(src + 16);
    const uint32_t w[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint32_t h[2];
      if constexpr (IL) {
        decode_u4x8_to_u8(w[i], zp4 ? 0x80808080u : 0u, h);
      } else {
        const uint32_t orv = zp4 ? 0x80808080u : 0u;
        const uint32_t t = lop3_and_or(w[i], 0x0f0f0f0fu, orv), u = lop3_and_or(w[i] >> 4, 0x0f0f0f0fu, orv);
        h[0] = __byte_perm(t, u, 0x5140); h[1] = __byte_perm(t, u, 0x7362);
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) out[2 * i + j] = zp4 ? bytes_sub_zp(h[j], zp4) : h[j];
    }
  }
}

// FMT: 0 = integer formats (LOP3 magic-number decode), 1 = 16-entry table formats (NF4 / fp4; BITS = 4, compressed storage),
//      2 = 8-bit float weights (e4m3 / e5m2; BITS = 8)
template <typename T, int BITS, int BM, bool IL, int FMT = 0>
__global__ void __launch_bounds__(TS_THREADS, BM <= 128 ? 2 : 1)
gemm_ts_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const TsParams p) {
  using SM = TsSmem<T, BITS, BM>;
  using EI = ElemInfo<T>;
  // the pipeline depth is a launch parameter only where two CTAs share an SM (BM <= 128: split-K / small grids); the
  // large-tile instance keeps it a compile-time constant (stage addressing folds into immediates: measured -12 % time)
  const int S = BM > 128 ? SM::kStages : p.stages;
  constexpr int KB = EI::kKB;
  constexpr int PRB = SM::kPRB;
  constexpr bool INT8 = EI::kInt8;
  constexpr int ACOL0 = TsTmem<BM>::kAcol0;
  constexpr int NCOLS = TsTmem<BM>::kCols;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sW = smem + S * SM::kActBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * SM::kStageBytes);
  uint64_t* full = bars;                 // [S]   TMA -> dequant + MMA
  uint64_t* empty = bars + S;            // [S]   MMA commit -> TMA
  uint64_t* a_ready = bars + 2 * S;      // [TA]  dequant -> MMA
  uint64_t* a_free = a_ready + TA_SLOTS; // [TA]  MMA commit -> dequant
  uint64_t* acc_full = a_free + TA_SLOTS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // split-K exists only for the small-M instances (BM <= 128); the large-tile instance keeps these compile-time constants
  constexpr bool SPLITK = BM <= 128;
  const int split = SPLITK ? blockIdx.x % p.splits : 0;
  const int tile = SPLITK ? blockIdx.x / p.splits : blockIdx.x;
  const int m_tile = tile % p.m_tiles;
  const int n_tile = tile / p.m_tiles;
  const int kb0 = SPLITK ? split * p.kb_per_split : 0;  // first k-block of this CTA
  const int m0 = m_tile * BM, n0 = n_tile * TS_ROWS;
  const int num_kb = SPLITK ? p.kb_per_split : p.K / KB;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    for (int i = 0; i < S; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < TA_SLOTS; ++i) { mbar_init(&a_ready[i], DQ_WARPS / dq_groups(BM)); mbar_init(&a_free[i], 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<NCOLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int s = 0;
      uint32_t par = 1;  // parity of the "previous use" of the stage: first pass must not wait
      for (int kb = 0; kb < num_kb; ++kb) {
        if (kb >= S) mbar_wait(&empty[s], par);
        mbar_arrive_expect_tx(&full[s], SM::kStageBytes);
        tma_load_2d(sA + s * SM::kActBytes, &tmA, (kb0 + kb) * KB, m0, &full[s]);
        if (p.w_tiled) {   // the same 128 x PRB bytes, gathered from four 32-row blocks of the slab-tiled storage
          const int byte = (kb0 + kb) * PRB;
          tma_load_4d(sW + s * SM::kWBytes, &tmW, byte % BB_TILE_ROW_BYTES, 0, byte / BB_TILE_ROW_BYTES, n0 / BB_TILE_ROWS, &full[s]);
        } else {
          tma_load_2d(sW + s * SM::kWBytes, &tmW, (kb0 + kb) * PRB, n0, &full[s]);
        }
        if (++s == S) { s = 0; par ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      uint32_t idesc;
      if constexpr (INT8) {
        const uint32_t a_signed = p.zp_const ? 1u : 0u;
        idesc = (2u << 4) | (a_signed << 7) | (1u << 10) | (uint32_t(BM >> 3) << 17) | (uint32_t(TS_ROWS >> 4) << 24);
      } else {
        const uint32_t f = std::is_same<T, __nv_bfloat16>::value ? 1u : 0u;
        idesc = (1u << 4) | (f << 7) | (f << 10) | (uint32_t(BM >> 3) << 17) | (uint32_t(TS_ROWS >> 4) << 24);
      }
      int s = 0, t = 0;
      uint32_t spar = 0, tpar = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full[s], spar);
        mbar_wait(&a_ready[t], tpar);
        tc_fence_after();
        const uint64_t bdesc = make_sw128_desc(smem_u32(sA + s * SM::kActBytes));
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          tc_mma_ts<INT8>(tmem_base, tmem_base + ACOL0 + t * 32 + kk * 8, bdesc + uint64_t(kk * 2), idesc,
                          (kb > 0 || kk > 0) ? 1u : 0u);
        }
        tc_commit(&empty[s]);
        tc_commit(&a_free[t]);
        if (++s == S) { s = 0; spar ^= 1; }
        if (++t == TsTmem<BM>::kSlots) { t = 0; tpar ^= 1; }
      }
      tc_commit(acc_full);
    }
  } else {
    // ===== dequant warps (then epilogue) =====
    const int quad = warp & 3;                         // TMEM lane quadrant this warp may touch
    constexpr int DQ_GROUPS = dq_groups(BM);
    constexpr int DQ_WARPS_PER_GROUP = DQ_WARPS / DQ_GROUPS;
    constexpr bool WHOLE_ROW = DQ_GROUPS == 4;          // this thread decodes both halves of its row
    const int half = WHOLE_ROW ? 0 : ((warp - 2) >> 2) & 1;   // which half of the k-block's columns (2-group organisation)
    const int grp = (warp - 2) / DQ_WARPS_PER_GROUP;   // which k-blocks (kb % DQ_GROUPS == grp)
    const int row = quad * 32 + lane;       // weight row inside the tile == TMEM lane
    const int n = n0 + row;
    const uint32_t lane_addr = tmem_base + (uint32_t(quad * 32) << 16);
    constexpr uint32_t MAGIC = INT8 ? 0u : TypeTraits<typename std::conditional<INT8, __half, T>::type>::kMagic;
    using TF = typename std::conditional<INT8, __half, T>::type;  // float-ish type for the 16-bit path

    constexpr bool HI = IL && !INT8 && std::is_same<TF, __half>::value && BITS == 4;
    constexpr uint32_t MAGIC_HI = HI ? 0x54005400u : MAGIC;
    constexpr int ZSH = HI ? 16 : 1;  // odd-nibble values are 64 + u: the folded zero point sits 4 mantissa bits up
    constexpr int NSLOT = TsTmem<BM>::kSlots;
    constexpr int SPG = NSLOT / DQ_GROUPS;   // operand slots per dequant group (1 or 2)
    static_assert(NSLOT <= TA_SLOTS && SPG * DQ_GROUPS == NSLOT && (SPG == 1 || SPG == 2), "slot ownership is static per group");
    const int kb_per_g = p.g / KB;
    const uint16_t* sc_row = reinterpret_cast<const uint16_t*>(p.scale) + size_t(n) * p.G;
    const uint16_t* z_row = reinterpret_cast<const uint16_t*>(p.zeros) + size_t(n) * p.G;
    constexpr int EPB = 8 / BITS;
    const uint8_t* qz_col = reinterpret_cast<const uint8_t*>(p.zeros) + n / EPB;
    const size_t qz_stride = size_t(p.N) * BITS / 8;
    const uint32_t qz_shift = BITS * (n % EPB);
    const uint32_t src0 = smem_u32(sW) + row * PRB + half * (PRB / 2);
    const uint32_t tdst0 = lane_addr + ACOL0 + half * 16;

    Lut4 lut4;
    if constexpr (FMT == 1) lut4_init(lut4, p.w_fmt, std::is_same<TF, __nv_bfloat16>::value, p.lut);
    auto dq_loop = [&](auto mode_tag) {
      constexpr int MODE = decltype(mode_tag)::value;
      DqConst c;
      c.mz_lo = MAGIC + uint32_t(p.zp_const) * 0x00010001u;
      c.mz_hi = MAGIC_HI + uint32_t(p.zp_const * ZSH) * 0x00010001u;
      c.z2 = c.s2 = c.negz2 = 0;
      // group parameters are fetched one needed-group ahead so their latency never sits in front of a decode
      uint16_t s_raw = 0, z_raw = 0;
      uint32_t qz_raw = 0;
      auto fetch_group = [&](int gi) {
        if constexpr (MODE != 0) s_raw = __ldg(sc_row + gi);
        if constexpr (MODE == 2 || MODE == 3) z_raw = __ldg(z_row + gi);
        if constexpr (MODE == 4) qz_raw = __ldg(qz_col + size_t(gi) * qz_stride);
      };
      int gi = (kb0 + grp) / kb_per_g;         // group of this warp's first k-block
      int g_end = (gi + 1) * kb_per_g - kb0;   // first (CTA-local) k-block of the next group
      bool fresh = true;
      if constexpr (MODE != 0) fetch_group(gi);
      int st = grp;                      // smem stage of kb (S % DQ_GROUPS == 0 keeps stage parity per group)
      uint32_t full_par = 0;
      uint32_t it = 0;                   // iteration count of this warp: slot = grp + DQ_GROUPS * (it mod SPG)
      int prev_slot = -1;
      const bool early_publish = S < DQ_GROUPS + 2;
      for (int kb = grp; kb < num_kb; kb += DQ_GROUPS) {
        if constexpr (MODE != 0) {
          if (kb >= g_end) {
            do { gi += 1; g_end += kb_per_g; } while (kb >= g_end);
            fresh = true;
          }
          if (fresh) {
            fresh = false;
            c.s2 = uint32_t(s_raw) * 0x00010001u;
            if constexpr (MODE == 2 || MODE == 3) {
              c.z2 = uint32_t(z_raw) * 0x00010001u;
              c.negz2 = c.z2 ^ 0x80008000u;
            }
            if constexpr (MODE == 4) {
              const uint32_t zq = (qz_raw >> qz_shift) & ((1u << BITS) - 1u);
              c.mz_lo = MAGIC + zq * 0x00010001u;
              c.mz_hi = MAGIC_HI + (zq * ZSH) * 0x00010001u;
            }
            // prefetch the parameters of the group this warp enters next: its first own k-block at or past g_end (the
            // stride DQ_GROUPS need not divide the group length -- g = 192 with four dequant groups skips 1 or 2 groups)
            const int kbn = kb + ((g_end - kb + DQ_GROUPS - 1) / DQ_GROUPS) * DQ_GROUPS;
            const int gn = gi + 1 + (kbn - g_end) / kb_per_g;
            if (kbn < num_kb && gn < p.G) fetch_group(gn);
          }
        }
        // The previous k-block's slot is normally published AFTER this block's first decode (its tcgen05.st hides behind
        // it).  With a pipeline no deeper than the group stride that order deadlocks: TMA(kb) waits for MMA(kb - S), which
        // waits for exactly the slot this warp would publish only after TMA(kb) has landed -- publish first then.
        if (early_publish && prev_slot >= 0) {
          tmem_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&a_ready[prev_slot]);
          prev_slot = -1;
        }
        mbar_wait(&full[st], full_par);
        uint32_t regs[16];
        if constexpr (INT8) dequant_half_row_i8<BITS, IL>(src0 + st * SM::kWBytes, regs, uint32_t(p.zp_const) * 0x01010101u);
        else if constexpr (FMT == 1) dequant_half_row_lut<TF, MODE>(src0 + st * SM::kWBytes, regs, lut4, c.s2);
        else if constexpr (FMT == 2) dequant_half_row_fp8<TF, MODE>(src0 + st * SM::kWBytes, regs, p.w_fmt == BB_W_FP8_E5M2, c.s2);
        else dequant_half_row<TF, BITS, MODE, IL>(src0 + st * SM::kWBytes, regs, c);
        // publish the PREVIOUS k-block's TMEM slot only now: its tcgen05.st had the whole decode above to land
        if (prev_slot >= 0) {
          tmem_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&a_ready[prev_slot]);
        }
        const int slot = grp + DQ_GROUPS * int(it & (SPG - 1));
        if (it >= SPG) {   // the MMAs that read this slot's previous contents must be done
          mbar_wait(&a_free[slot], ((it / SPG) - 1) & 1);
          tc_fence_after();
        }
        tmem_st_x16(tdst0 + slot * 32, regs);
        if constexpr (WHOLE_ROW) {
          if constexpr (INT8) dequant_half_row_i8<BITS, IL>(src0 + st * SM::kWBytes + PRB / 2, regs, uint32_t(p.zp_const) * 0x01010101u);
          else if constexpr (FMT == 1) dequant_half_row_lut<TF, MODE>(src0 + st * SM::kWBytes + PRB / 2, regs, lut4, c.s2);
          else if constexpr (FMT == 2) dequant_half_row_fp8<TF, MODE>(src0 + st * SM::kWBytes + PRB / 2, regs, p.w_fmt == BB_W_FP8_E5M2, c.s2);
          else dequant_half_row<TF, BITS, MODE, IL>(src0 + st * SM::kWBytes + PRB / 2, regs, c);
          tmem_st_x16(tdst0 + slot * 32 + 16, regs);
        }
        prev_slot = slot;
        ++it;
        st += DQ_GROUPS;
        if (st >= S) { st -= S; full_par ^= 1; }
      }
      if (prev_slot >= 0) {
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_ready[prev_slot]);
      }
    };
    if constexpr (INT8) {
      dq_loop(std::integral_constant<int, 0>{});
    } else if constexpr (FMT != 0) {   // table / fp8 formats: no zero points on the fast path (gemm_ts_supported)
      if (p.mode == 0) dq_loop(std::integral_constant<int, 0>{}); else dq_loop(std::integral_constant<int, 1>{});
    } else {
      switch (p.mode) {
        case 0: dq_loop(std::integral_constant<int, 0>{}); break;
        case 1: dq_loop(std::integral_constant<int, 1>{}); break;
        case 2: dq_loop(std::integral_constant<int, 2>{}); break;
        case 3: dq_loop(std::integral_constant<int, 3>{}); break;
        default: dq_loop(std::integral_constant<int, 4>{}); break;
      }
    }

    // ----- epilogue: accumulator[lane = n, column = m] -> C[m, n] -----
    mbar_wait(acc_full, 0);
    tc_fence_after();
    constexpr int CPH = BM / 4;                  // accumulator columns per warp (4 warps share a lane quadrant)
    constexpr int CH = CPH < 16 ? CPH : 16;      // columns per tcgen05.ld
    static_assert(CPH % CH == 0 && (CH == 16 || CH == 8), "BM must be >= 32");
    float bias_f = 0.f;
    if (p.bias) {
      if (p.a_dtype == BB_F16) bias_f = __half2float(reinterpret_cast<const __half*>(p.bias)[n]);
      else if (p.a_dtype == BB_BF16) bias_f = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.bias)[n]);
      else bias_f = float(reinterpret_cast<const int8_t*>(p.bias)[n]);
    }
    const int combo = WHOLE_ROW ? grp : grp * 2 + half;   // the four quadrant warps with equal `combo` own the same CPH columns
    const int cbase = combo * CPH;
    void* const C0 = p.out.ptr[0];
    const int ndst = p.out.n;
    const size_t ld = size_t(p.out.ld), col0 = size_t(p.out.col0);
    if (!INT8 && p.staged_epi && !(SPLITK && p.splits > 1)) {
      // Staged epilogue (fp16 / bf16 outputs).  The accumulator is [lane = n][column = m]; stored straight from registers a
      // warp instruction covers 32 consecutive n of ONE row m with 2-byte elements -- 64 bytes per instruction and, on the
      // column-parallel path, per peer: 8 destinations x 2-byte stores over NVLink (round-1 verdict: 4 -> 8 GPUs bought 7 %).
      // Here the four quadrant warps of a column range transpose their [128 n] x [CPH m] tile through the (now idle) pipeline
      // smem and write 256-byte row segments with 16-byte stores: 1/8 of the store instructions, full 128-byte lines per peer.
      uint8_t* tile = smem + combo * (CPH * 256);
      const uint32_t tcol = smem_u32(tile) + uint32_t(quad * 32 + lane) * 2u;
      for (int c0 = cbase; c0 < cbase + CPH; c0 += CH) {
        if (m0 + c0 >= p.M) break;  // uniform over the four warps
        uint32_t v[16];
        if constexpr (CH == 16) tmem_ld_x16(lane_addr + c0, v); else tmem_ld_x8(lane_addr + c0, v);
        tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const float acc = __uint_as_float(v[c]);
          uint16_t bits;
          if (p.out_dtype == BB_F16) {
            __half h = __float2half_rn(acc);
            if (p.bias) h = __hadd(h, __float2half_rn(bias_f));
            bits = __half_as_ushort(h);
          } else {
            __nv_bfloat16 h = __float2bfloat16_rn(acc);
            if (p.bias) h = __hadd(h, __float2bfloat16_rn(bias_f));
            bits = __bfloat16_as_ushort(h);
          }
          asm volatile("st.shared.u16 [%0], %1;" ::"r"(tcol + uint32_t(c0 - cbase + c) * 256u), "h"(bits) : "memory");
        }
      }
      asm volatile("bar.sync %0, %1;" ::"r"(1 + combo), "n"(128) : "memory");
      const int rows_valid = min(CPH, p.M - (m0 + cbase));
      const int t128 = quad * 32 + lane;
#pragma unroll 1
      for (int id = t128; id < CPH * 16; id += 128) {
        const int row = id >> 4, c16 = id & 15;
        if (row >= rows_valid) break;
        uint4 q;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w)
                     : "r"(smem_u32(tile) + uint32_t(row) * 256u + uint32_t(c16) * 16u));
        const size_t o = (size_t(m0 + cbase + row) * ld + col0 + size_t(n0) + size_t(c16) * 8u) * 2u;
        *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(C0) + o) = q;
        for (int d = 1; d < ndst; ++d) *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(p.out.ptr[d]) + o) = q;
      }
    } else
    for (int c0 = cbase; c0 < cbase + CPH; c0 += CH) {
      if (m0 + c0 >= p.M) break;  // warp-uniform
      uint32_t v[16];
      if constexpr (CH == 16) tmem_ld_x16(lane_addr + c0, v); else tmem_ld_x8(lane_addr + c0, v);
      tmem_wait_ld();
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const int m = m0 + c0 + c;
        if (m >= p.M) break;
        if (SPLITK && p.splits > 1) {  // raw partial accumulator; bias / cast / scatter happen in splitk_reduce_kernel
          reinterpret_cast<uint32_t*>(p.ws)[(size_t(split) * p.M + m) * p.N + n] = v[c];
          continue;
        }
        const size_t o = size_t(m) * ld + col0 + size_t(n);
        // destination 0 (the local output, or this rank's own copy) takes the straight-line path the single-GPU kernel
        // always had; the peers' copies (column-parallel scatter, NVLink) follow in a loop that is empty otherwise --
        // indexing the pointer table per element cost ~9 % of the M=4096 GEMM
        if constexpr (INT8) {
          const int acc = int(v[c]);
          const int b = int(bias_f);
          auto put = [&](void* Cd) {
            switch (p.out_dtype) {
              case BB_I32: reinterpret_cast<int*>(Cd)[o] = acc + b; break;
              case BB_I8: reinterpret_cast<int8_t*>(Cd)[o] = int8_t(int8_t(acc) + b); break;
              case BB_F32: reinterpret_cast<float*>(Cd)[o] = float(acc) + float(b); break;
              case BB_F16: reinterpret_cast<__half*>(Cd)[o] = __hadd(__int2half_rn(acc), __int2half_rn(b)); break;
              default: reinterpret_cast<__nv_bfloat16*>(Cd)[o] = __hadd(__int2bfloat16_rn(acc), __int2bfloat16_rn(b));
            }
          };
          put(C0);
          for (int d = 1; d < ndst; ++d) put(p.out.ptr[d]);
        } else {
          const float acc = __uint_as_float(v[c]);
          if (p.out_dtype == BB_F16) {
            __half h = __float2half_rn(acc);
            if (p.bias) h = __hadd(h, __float2half_rn(bias_f));
            reinterpret_cast<__half*>(C0)[o] = h;
            for (int d = 1; d < ndst; ++d) reinterpret_cast<__half*>(p.out.ptr[d])[o] = h;
          } else if (p.out_dtype == BB_BF16) {
            __nv_bfloat16 h = __float2bfloat16_rn(acc);
            if (p.bias) h = __hadd(h, __float2bfloat16_rn(bias_f));
            reinterpret_cast<__nv_bfloat16*>(C0)[o] = h;
            for (int d = 1; d < ndst; ++d) reinterpret_cast<__nv_bfloat16*>(p.out.ptr[d])[o] = h;
          } else {
            const float f = acc + (p.bias ? bias_f : 0.f);
            reinterpret_cast<float*>(C0)[o] = f;
            for (int d = 1; d < ndst; ++d) reinterpret_cast<float*>(p.out.ptr[d])[o] = f;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<NCOLS>(tmem_base);
  }
}

// out[m, n] = cast(sum_s ws[s][m][n]) (+ bias), stored to every destination of the OutSpec.  One thread = 4 consecutive n
// (N is a multiple of 128 here): 128-bit loads of the partials, 32-bit index arithmetic, one vector store per destination.
template <bool INT8>
__global__ void splitk_reduce_kernel(const TsParams p) {
  const unsigned n4 = unsigned(p.N) >> 2;                 // quads per row
  const unsigned total4 = unsigned(p.M) * n4;             // m * N / 4 < 2^31 (m <= 128 on this path)
  const size_t split_stride4 = size_t(p.M) * n4;          // quads per split
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += gridDim.x * blockDim.x) {
    const unsigned m = i / n4, n = (i - m * n4) * 4u;
    float bias_f[4] = {0.f, 0.f, 0.f, 0.f};
    if (p.bias) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (p.a_dtype == BB_F16) bias_f[j] = __half2float(reinterpret_cast<const __half*>(p.bias)[n + j]);
        else if (p.a_dtype == BB_BF16) bias_f[j] = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.bias)[n + j]);
        else bias_f[j] = float(reinterpret_cast<const int8_t*>(p.bias)[n + j]);
      }
    }
    const size_t o = size_t(m) * size_t(p.out.ld) + size_t(p.out.col0) + n;
    if constexpr (INT8) {
      int acc[4] = {0, 0, 0, 0};
      for (int sp = 0; sp < p.splits; ++sp) {
        const int4 v = __ldcs(reinterpret_cast<const int4*>(p.ws) + size_t(sp) * split_stride4 + i);
        acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
      }
      for (int d = 0; d < p.out.n; ++d) {
        void* Cd = p.out.ptr[d];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int b = int(bias_f[j]);
          switch (p.out_dtype) {
            case BB_I32: reinterpret_cast<int*>(Cd)[o + j] = acc[j] + b; break;
            case BB_I8: reinterpret_cast<int8_t*>(Cd)[o + j] = int8_t(int8_t(acc[j]) + b); break;
            case BB_F32: reinterpret_cast<float*>(Cd)[o + j] = float(acc[j]) + float(b); break;
            case BB_F16: reinterpret_cast<__half*>(Cd)[o + j] = __hadd(__int2half_rn(acc[j]), __int2half_rn(b)); break;
            default: reinterpret_cast<__nv_bfloat16*>(Cd)[o + j] = __hadd(__int2bfloat16_rn(acc[j]), __int2bfloat16_rn(b));
          }
        }
      }
    } else {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      for (int sp = 0; sp < p.splits; ++sp) {
        const float4 v = __ldcs(reinterpret_cast<const float4*>(p.ws) + size_t(sp) * split_stride4 + i);
        acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
      }
      if (p.out_dtype == BB_F16) {
        __half h[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          h[j] = __float2half_rn(acc[j]);
          if (p.bias) h[j] = __hadd(h[j], __float2half_rn(bias_f[j]));
        }
        for (int d = 0; d < p.out.n; ++d) {
          __half* dst = reinterpret_cast<__half*>(p.out.ptr[d]) + o;
#pragma unroll
          for (int j = 0; j < 4; ++j) dst[j] = h[j];
        }
      } else if (p.out_dtype == BB_BF16) {
        __nv_bfloat16 h[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          h[j] = __float2bfloat16_rn(acc[j]);
          if (p.bias) h[j] = __hadd(h[j], __float2bfloat16_rn(bias_f[j]));
        }
        for (int d = 0; d < p.out.n; ++d) {
          __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out.ptr[d]) + o;
#pragma unroll
          for (int j = 0; j < 4; ++j) dst[j] = h[j];
        }
      } else {
        for (int d = 0; d < p.out.n; ++d) {
          float* dst = reinterpret_cast<float*>(p.out.ptr[d]) + o;
#pragma unroll
          for (int j = 0; j < 4; ++j) dst[j] = acc[j] + (p.bias ? bias_f[j] : 0.f);
        }
      }
    }
  }
}

// split-K factor: only when the tile grid leaves most SMs idle; the k-range must split on k-block boundaries
int choose_splits(int tiles, int num_kb, int sms) {
  if (tiles >= sms) return 1;
  int best = 1;
  for (int sp = 2; sp <= 8; ++sp) {
    if (num_kb % sp || num_kb / sp < 8) continue;
    if (tiles * sp <= 2 * sms) best = sp;  // two CTAs per SM are resident for BM <= 128
  }
  return best;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
std::once_flag g_encode_once;

EncodeTiledFn get_encode() {
  std::call_once(g_encode_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  });
  return g_encode;
}

int make_map_2d(CUtensorMap* map, CUtensorMapDataType dt, const void* base, uint64_t inner, uint64_t outer,
                uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle sw) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return 4; }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", int(r)); return 4; }
  return 0;
}

// BB_TILE_SLAB storage [N/32][row_bytes/512][32][512 B] as a 4-D tensor; box = {PRB bytes, 32 rows, 1 segment, 4 row blocks}
// lands in shared memory as the same dense [128 rows][PRB] tile the row-major 2-D box produces
int make_map_w_tiled(CUtensorMap* map, const void* base, uint64_t row_bytes, uint64_t rows, uint32_t prb) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return 4; }
  const uint64_t upr = row_bytes / BB_TILE_ROW_BYTES;
  cuuint64_t dims[4] = {BB_TILE_ROW_BYTES, BB_TILE_ROWS, upr, rows / BB_TILE_ROWS};
  cuuint64_t strides[3] = {BB_TILE_ROW_BYTES, uint64_t(BB_TILE_ROW_BYTES) * BB_TILE_ROWS, uint64_t(BB_TILE_ROW_BYTES) * BB_TILE_ROWS * upr};
  cuuint32_t box[4] = {prb, BB_TILE_ROWS, 1, TS_ROWS / BB_TILE_ROWS};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (4-D, slab-tiled W) failed with CUresult %d", int(r)); return 4; }
  return 0;
}

template <typename T, int BITS, int BM, bool IL, int FMT = 0>
int launch_ts_inst(const MatmulArgs& a, const TsParams& p0) {
  using SM = TsSmem<T, BITS, BM>;
  using EI = ElemInfo<T>;
  auto kernel = gemm_ts_kernel<T, BITS, BM, IL, FMT>;
  static bool attr_set[BB_MAX_DEVICES] = {};   // cudaFuncSetAttribute is per device
  const int dev = current_device();
  if (!attr_set[dev]) {
    BB_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kTotal));
    attr_set[dev] = true;
  }
  TsParams p = p0;
  p.m_tiles = (a.m + BM - 1) / BM;
  const int num_kb_total = a.d.K / EI::kKB;
  const int tiles = p.m_tiles * (a.d.N / TS_ROWS);
  p.splits = BM <= 128 ? choose_splits(tiles, num_kb_total, device_sm_count()) : 1;
  if (p.splits > 1 && a.workspace_bytes < size_t(p.splits) * a.m * a.d.N * 4) p.splits = 1;  // no scratch given
  p.kb_per_split = num_kb_total / p.splits;
  p.ws = a.workspace;
  // two CTAs per SM when the kernel is split (or simply small): cap the pipeline so that 2 x smem fits
  int stages = SM::kStages;
  if (BM <= 128 && tiles * p.splits > device_sm_count()) {
    // more CTAs than SMs: keep the pipeline shallow enough for two resident CTAs (their windows add up); with at most
    // one CTA per SM the full depth is needed to cover the TMA latency
    const int cap = ((110 * 1024 - SM::kBarBytes - 1024) / SM::kStageBytes) & ~1;
    if (cap >= dq_groups(BM) && cap < stages) stages = cap;   // (a dequant group's first stage index is its group id)
  }
  p.stages = stages;
  {
    // staged epilogue: default on the column-parallel (multi-destination) path; BB_TS_STAGED=0/1 forces it.  Needs 16-byte
    // aligned row segments in every destination and BM x 256 B of (idle) pipeline smem
    static const int force = [] { const char* e = getenv("BB_TS_STAGED"); return e ? atoi(e) : -1; }();
    const OutSpec& o = p.out;
    bool ok = !EI::kInt8 && (a.d.out_dtype == BB_F16 || a.d.out_dtype == BB_BF16) && (o.ld % 8) == 0 && (o.col0 % 8) == 0 &&
              size_t(stages) * SM::kStageBytes >= size_t(BM) * 256;
    for (int i = 0; i < o.n && ok; ++i) ok = (reinterpret_cast<uintptr_t>(o.ptr[i]) & 15) == 0;
    p.staged_epi = ok && (force < 0 ? o.n > 1 : force != 0) ? 1 : 0;
  }
  const size_t smem_bytes = size_t(stages) * SM::kStageBytes + SM::kBarBytes + 1024;
  CUtensorMap tmA, tmW;
  const CUtensorMapDataType adt = EI::kInt8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                                            : (std::is_same<T, __half>::value ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                                                               : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
  int rc = make_map_2d(&tmA, adt, a.A, uint64_t(a.d.K), uint64_t(a.m), uint64_t(a.d.K) * sizeof(T), EI::kKB, BM,
                       CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  const uint64_t wrow = uint64_t(a.d.K) * BITS / 8;
  p.w_tiled = a.d.w_tile == BB_TILE_SLAB ? 1 : 0;
  if (p.w_tiled) rc = make_map_w_tiled(&tmW, a.W, wrow, uint64_t(a.d.N), SM::kPRB);
  else rc = make_map_2d(&tmW, CU_TENSOR_MAP_DATA_TYPE_UINT8, a.W, wrow, uint64_t(a.d.N), wrow, SM::kPRB, TS_ROWS,
                        CU_TENSOR_MAP_SWIZZLE_NONE);
  if (rc) return rc;
  const int grid = tiles * p.splits;
  kernel<<<grid, TS_THREADS, smem_bytes, a.stream>>>(tmA, tmW, p);
  BB_LAUNCH_CHECK();
  if (p.splits > 1) {
    const size_t total = size_t(a.m) * a.d.N / 4;   // one thread per 4 consecutive n
    const int rthreads = 256;
    const int rblocks = int(std::min<size_t>((total + rthreads - 1) / rthreads, size_t(device_sm_count()) * 8));
    if (EI::kInt8) splitk_reduce_kernel<true><<<rblocks, rthreads, 0, a.stream>>>(p);
    else splitk_reduce_kernel<false><<<rblocks, rthreads, 0, a.stream>>>(p);
    BB_LAUNCH_CHECK();
  }
  return 0;
}

template <typename T, int BITS, bool IL>
int launch_ts_bm2(const MatmulArgs& a, const TsParams& p) {
  if (a.m <= 32) return launch_ts_inst<T, BITS, 32, IL>(a, p);
  if (a.m <= 64) return launch_ts_inst<T, BITS, 64, IL>(a, p);
  if (a.m <= 128) return launch_ts_inst<T, BITS, 128, IL>(a, p);
  // tuning hook: BB_TS_BM=128 runs large m on the 128-column tile (two CTAs per SM: one CTA's epilogue overlaps the other's main loop,
  // at twice the decode work per output)
  static const int force_bm = [] { const char* e = getenv("BB_TS_BM"); return e ? atoi(e) : 0; }();
  if (force_bm == 128) return launch_ts_inst<T, BITS, 128, IL>(a, p);
  return launch_ts_inst<T, BITS, 256, IL>(a, p);
}
template <typename T>
int launch_ts_lut(const MatmulArgs& a, const TsParams& p) {
  if (a.m <= 32) return launch_ts_inst<T, 4, 32, false, 1>(a, p);
  if (a.m <= 64) return launch_ts_inst<T, 4, 64, false, 1>(a, p);
  if (a.m <= 128) return launch_ts_inst<T, 4, 128, false, 1>(a, p);
  return launch_ts_inst<T, 4, 256, false, 1>(a, p);
}
template <typename T>
int launch_ts_fp8(const MatmulArgs& a, const TsParams& p) {
  if (a.m <= 32) return launch_ts_inst<T, 8, 32, false, 2>(a, p);
  if (a.m <= 64) return launch_ts_inst<T, 8, 64, false, 2>(a, p);
  if (a.m <= 128) return launch_ts_inst<T, 8, 128, false, 2>(a, p);
  return launch_ts_inst<T, 8, 256, false, 2>(a, p);
}
template <typename T, int BITS>
int launch_ts_bm(const MatmulArgs& a, const TsParams& p) {
  return a.d.w_layout == BB_LAYOUT_COMPRESSED ? launch_ts_bm2<T, BITS, false>(a, p) : launch_ts_bm2<T, BITS, true>(a, p);
}

}  // namespace

bool gemm_ts_supported(const bb_matmul_desc& d, int m) {
  if (m < 1) return false;
  const bool table_fmt = d.w_fmt == BB_W_NF || d.w_fmt == BB_W_FP4;
  if (table_fmt) {   // 16-entry table formats: 4-bit compressed storage, scale only (a zero point has no fast path)
    if (d.w_bits != 4 || d.w_layout != BB_LAYOUT_COMPRESSED || d.with_zeros || d.a_dtype == BB_I8) return false;
  } else if (d.w_fmt == BB_W_FP8_E4M3 || d.w_fmt == BB_W_FP8_E5M2) {   // 8-bit float weights x 16-bit activations, scale only
    if (d.w_bits != 8 || d.w_layout != BB_LAYOUT_COMPRESSED || d.with_zeros || d.a_dtype == BB_I8) return false;
  } else if (d.w_fmt != BB_W_UINT && d.w_fmt != BB_W_INT) {
    return false;
  }
  if (d.w_bits != 4 && d.w_bits != 2 && !(d.w_bits == 8 && (d.w_fmt == BB_W_FP8_E4M3 || d.w_fmt == BB_W_FP8_E5M2))) return false;
  if (d.N % TS_ROWS) return false;
  const int g = d.group_size <= 0 ? d.K : d.group_size;
  if (d.a_dtype == BB_I8) {
    if (d.accum_dtype != BB_I32 || d.w_layout == BB_LAYOUT_INTERLEAVED_16) return false;
    if (d.with_scaling || d.with_zeros) return false;
    if (d.K % 128) return false;
    return true;
  }
  if (d.a_dtype != BB_F16 && d.a_dtype != BB_BF16) return false;
  if (d.w_layout == BB_LAYOUT_INTERLEAVED_8) return false;
  if (d.K % 64 || g % 64 || d.K % g) return false;
  if (d.with_zeros && !d.with_scaling) return false;
  if (d.w_fmt == BB_W_INT && d.with_zeros) return false;
  if (d.out_dtype != BB_F16 && d.out_dtype != BB_BF16 && d.out_dtype != BB_F32) return false;
  if (d.with_zeros && d.zeros_mode == BB_ZEROS_QUANTIZED && (d.N * d.w_bits) % 8) return false;
  return true;
}

size_t gemm_ts_workspace_bytes(const bb_matmul_desc& d, int m) {
  if (m > 128) return 0;
  const int kb = d.a_dtype == BB_I8 ? 128 : 64;
  const int bm = m <= 32 ? 32 : (m <= 64 ? 64 : 128);
  const int tiles = ((m + bm - 1) / bm) * (d.N / TS_ROWS);
  const int sp = choose_splits(tiles, d.K / kb, device_sm_count());
  return sp > 1 ? size_t(sp) * m * d.N * 4 : 0;
}

int gemm_ts_init(int) { return get_encode() ? 0 : 0; }

int launch_gemm_ts(const MatmulArgs& a) {
  const bb_matmul_desc& d = a.d;
  if ((reinterpret_cast<uintptr_t>(a.A) & 15) || (reinterpret_cast<uintptr_t>(a.W) & 15)) {
    set_error("gemm_ts: A and W must be 16-byte aligned");
    return 5;
  }
  TsParams p;
  p.scale = d.with_scaling ? a.scale : nullptr;
  p.zeros = d.with_zeros ? a.zeros : nullptr;
  p.bias = d.with_bias ? a.bias : nullptr;
  p.out = make_outspec(a); p.M = a.m; p.N = d.N; p.K = d.K; p.g = a.gsize(); p.G = a.groups();
  p.mode = !d.with_scaling ? 0 : (!d.with_zeros ? 1 : 2 + d.zeros_mode);
  p.zp_const = d.w_fmt == BB_W_INT ? (1 << (d.w_bits - 1)) : 0;
  p.out_dtype = d.out_dtype; p.a_dtype = d.a_dtype; p.m_tiles = 1;
  p.lut = d.w_fmt == BB_W_NF ? a.lut : nullptr;
  p.w_fmt = d.w_fmt;
  if (d.w_fmt == BB_W_NF || d.w_fmt == BB_W_FP4) return d.a_dtype == BB_F16 ? launch_ts_lut<__half>(a, p) : launch_ts_lut<__nv_bfloat16>(a, p);
  if (d.w_fmt == BB_W_FP8_E4M3 || d.w_fmt == BB_W_FP8_E5M2) return d.a_dtype == BB_F16 ? launch_ts_fp8<__half>(a, p) : launch_ts_fp8<__nv_bfloat16>(a, p);
  if (d.a_dtype == BB_I8) return d.w_bits == 4 ? launch_ts_bm<int8_t, 4>(a, p) : launch_ts_bm<int8_t, 2>(a, p);
  if (d.a_dtype == BB_F16) return d.w_bits == 4 ? launch_ts_bm<__half, 4>(a, p) : launch_ts_bm<__half, 2>(a, p);
  return d.w_bits == 4 ? launch_ts_bm<__nv_bfloat16, 4>(a, p) : launch_ts_bm<__nv_bfloat16, 2>(a, p);
}

}  // namespace bb
