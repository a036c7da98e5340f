// bb_prep.cu -- weight pre-processing and ingest (one-off, but 70B-parameter models make it matter).
//
// Replaces the reference's TVM-LLVM CPU ops QuantCompress (bitblas/ops/quant_compress/quant_compress_impl.py:22-30)
// and LOP3Permutate (bitblas/ops/lop3_permutate/lop3_permutate_impl.py:27-34), chained on the host by
// OPExecutorCPU (bitblas/ops/operator.py:529-556), and the Python column loops of the GPTQ repack
// (bitblas/module/__init__.py:24-74,315-363) with single-pass device kernels (one 32-bit output word per
// thread) plus plain C++ host versions.
#include "bb_common.cuh"

namespace bb {
namespace {

__global__ void transform_weight_kernel(const int8_t* __restrict__ w, uint32_t* __restrict__ out, int64_t rows,
                                        int64_t cols, int bits, int layout) {
  const int epw = 32 / bits;
  const int64_t words_per_row = cols / epw;
  const int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (idx >= rows * words_per_row) return;
  const int64_t rrow = idx / words_per_row, kw = idx % words_per_row;
  const int8_t* src = w + rrow * cols + kw * epw;
  const uint32_t mask = (1u << bits) - 1u;
  uint32_t word = 0;
  for (int o = 0; o < epw; ++o) word |= (uint32_t(uint8_t(src[o])) & mask) << field_bitpos(o, bits, layout);
  out[idx] = word;
}

// GPTQ qweight int32 [K/epw, N] -> BitBLAS [N, K/epw] words with the layout's bit permutation
__global__ void repack_gptq_qweight_kernel(const uint32_t* __restrict__ q, uint32_t* __restrict__ out, int64_t KW,
                                           int64_t N, int bits, int layout) {
  __shared__ uint32_t tile[32][33];
  const int64_t n0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t kw = k0 + i, n = n0 + threadIdx.x;
    tile[i][threadIdx.x] = (kw < KW && n < N) ? q[kw * N + n] : 0u;
  }
  __syncthreads();
  const int epw = 32 / bits;
  const uint32_t mask = (1u << bits) - 1u;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t n = n0 + i, kw = k0 + threadIdx.x;
    if (n >= N || kw >= KW) continue;
    const uint32_t src = tile[threadIdx.x][i];
    uint32_t word = 0;
    for (int o = 0; o < epw; ++o) word |= ((src >> (bits * o)) & mask) << field_bitpos(o, bits, layout);
    out[n * KW + kw] = word;
  }
}

template <typename T>
__global__ void repack_gptq_qzeros_kernel(const uint32_t* __restrict__ qz, const T* __restrict__ scales, void* out,
                                          int64_t G, int64_t N, int bits, int zeros_mode, int v2) {
  const int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  const int epw = 32 / bits;
  const uint32_t mask = (1u << bits) - 1u;
  if (zeros_mode == BB_ZEROS_QUANTIZED) {
    // one output byte per thread: out[G, N*bits/8]
    const int epb = 8 / bits;
    const int64_t bytes_per_row = N / epb;
    if (idx >= G * bytes_per_row) return;
    const int64_t gi = idx / bytes_per_row, nb = idx % bytes_per_row;
    uint32_t byte = 0;
    for (int e = 0; e < epb; ++e) {
      const int64_t n = nb * epb + e;
      uint32_t z = (qz[gi * (N / epw) + n / epw] >> (bits * (n % epw))) & mask;
      if (!v2) z = (z + 1) & mask;
      byte |= z << (bits * e);
    }
    reinterpret_cast<uint8_t*>(out)[idx] = uint8_t(byte);
  } else {
    if (idx >= G * N) return;
    const int64_t n = idx / G, gi = idx % G;  // out[N, G]
    uint32_t z = (qz[gi * (N / epw) + n / epw] >> (bits * (n % epw))) & mask;
    if (!v2) z = (z + 1) & mask;
    T zv = TypeTraits<T>::from_float(float(z));
    if (zeros_mode == BB_ZEROS_RESCALE) zv = __hmul(zv, scales[idx]);
    reinterpret_cast<T*>(out)[idx] = zv;
  }
}

template <typename T>
__global__ void debug_decode16_kernel(int bits, int zp, int layout, const uint32_t* __restrict__ in, T* __restrict__ out, int nwords) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nwords) return;
  const uint32_t w = in[i];
  const uint32_t mz = TypeTraits<T>::kMagic + uint32_t(zp) * 0x00010001u;
  if (bits == 4) {
    uint32_t h[4];
    decode_u4x8_raw<T>(w, h);
    for (int j = 0; j < 4; ++j) {
      const uint32_t v = sub2<T>(h[j], mz);
      const uint16_t lo = uint16_t(v & 0xffff), hi = uint16_t(v >> 16);
      // interleaved: (u[2j], u[2j+1]); compressed: (u[j], u[j+4])
      const int i0 = layout == BB_LAYOUT_COMPRESSED ? j : 2 * j, i1 = layout == BB_LAYOUT_COMPRESSED ? j + 4 : 2 * j + 1;
      out[8 * i + i0] = *reinterpret_cast<const T*>(&lo);
      out[8 * i + i1] = *reinterpret_cast<const T*>(&hi);
    }
  } else {
    uint32_t h[8];
    if (layout == BB_LAYOUT_COMPRESSED) decode_u2x16_raw_compressed<T>(w, h); else decode_u2x16_raw_interleaved<T>(w, h);
    for (int j = 0; j < 8; ++j) {
      const uint32_t v = sub2<T>(h[j], mz);
      const uint16_t lo = uint16_t(v & 0xffff), hi = uint16_t(v >> 16);
      const int i0 = layout == BB_LAYOUT_COMPRESSED ? j : 2 * j, i1 = layout == BB_LAYOUT_COMPRESSED ? j + 8 : 2 * j + 1;
      out[16 * i + i0] = *reinterpret_cast<const T*>(&lo);
      out[16 * i + i1] = *reinterpret_cast<const T*>(&hi);
    }
  }
}

// the GEMM path's dequant arithmetic (dq_finish) over an array of packed words; one (scale, zeros, qzeros) triple per group of 8
// outputs, like the reference's decode_*_scale[_zeros_*] device functions (fast_decoding.hpp) which the KATs compare against
template <typename T, int MODE>
__global__ void debug_dequant16_kernel(int bits, int zp, int layout, const uint32_t* __restrict__ in, const T* __restrict__ scale,
                                       const T* __restrict__ zeros, const int* __restrict__ qzeros, T* __restrict__ out, int nwords) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nwords) return;
  const uint32_t w = in[i];
  const int per_word = 32 / bits, gpw = per_word / 8;   // outputs and 8-output groups per 32-bit word
  uint32_t h[8];
  if (bits == 4) { uint32_t t[4]; decode_u4x8_raw<T>(w, t); for (int j = 0; j < 4; ++j) h[j] = t[j]; }
  else if (layout == BB_LAYOUT_COMPRESSED) decode_u2x16_raw_compressed<T>(w, h);
  else decode_u2x16_raw_interleaved<T>(w, h);
  for (int j = 0; j < per_word / 2; ++j) {
    const int i0 = layout == BB_LAYOUT_COMPRESSED ? j : 2 * j, i1 = layout == BB_LAYOUT_COMPRESSED ? j + per_word / 2 : 2 * j + 1;
    uint16_t r[2];
    const int idx[2] = {i0, i1};
    for (int e = 0; e < 2; ++e) {   // the two halves of a pair may belong to different 8-output groups (compressed layout)
      const int g = i * gpw + idx[e] / 8;
      DqConst c;
      const uint32_t zq = (MODE == 4) ? uint32_t(qzeros[g]) : 0u;
      c.mz_lo = c.mz_hi = TypeTraits<T>::kMagic + (uint32_t(zp) + zq) * 0x00010001u;
      c.s2 = MODE != 0 ? dup2<T>(scale[g]) : 0u;
      c.z2 = (MODE == 2 || MODE == 3) ? dup2<T>(zeros[g]) : 0u;
      c.negz2 = c.z2 ^ 0x80008000u;
      const uint32_t v = bits == 2 ? dq_finish<T, MODE, 2>(h[j], c.mz_lo, c) : dq_finish<T, MODE, 4>(h[j], c.mz_lo, c);
      r[e] = e == 0 ? uint16_t(v & 0xffff) : uint16_t(v >> 16);
    }
    out[size_t(per_word) * i + i0] = *reinterpret_cast<const T*>(&r[0]);
    out[size_t(per_word) * i + i1] = *reinterpret_cast<const T*>(&r[1]);
  }
}

__global__ void debug_decode8_kernel(int bits, int zp, const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int nwords) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nwords) return;
  const uint32_t w = in[i];
  const uint32_t zp4 = uint32_t(zp) * 0x01010101u;
  if (bits == 2) {
    uint32_t h[4];
    decode_u2x16_to_u8(w, zp ? 0x80808080u : 0u, h);
    for (int j = 0; j < 4; ++j) out[4 * i + j] = zp ? bytes_sub_zp(h[j], zp4) : h[j];
  } else {
    uint32_t h[2];
    decode_u4x8_to_u8(w, zp ? 0x80808080u : 0u, h);
    for (int j = 0; j < 2; ++j) out[2 * i + j] = zp ? bytes_sub_zp(h[j], zp4) : h[j];
  }
}

int layout_from_target(int target_bits) {
  return target_bits == 0 ? BB_LAYOUT_COMPRESSED : (target_bits == 8 ? BB_LAYOUT_INTERLEAVED_8 : BB_LAYOUT_INTERLEAVED_16);
}

}  // namespace
}  // namespace bb

using namespace bb;

// BB_TILE_SLAB re-tiling: one 16-byte chunk per thread; consecutive threads walk the ROW-MAJOR side (coalesced reads going in,
// coalesced writes coming back), the tiled side is contiguous per 512-byte segment
static __global__ void retile_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int64_t rows, int64_t row_bytes, int inverse) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t cpr = row_bytes / 16;
  if (i >= rows * cpr) return;
  const int64_t n = i / cpr, b = (i - n * cpr) * 16;
  const size_t t = tiled_byte_offset(n, b, row_bytes) / 16;
  if (inverse) out[i] = in[t]; else out[t] = in[i];
}

extern "C" {

int bb_compress_host(const int8_t* in, int8_t* out, int64_t rows, int64_t cols, int bits) {
  if (!in || !out || (bits != 1 && bits != 2 && bits != 4) || cols % (8 / bits)) { set_error("bb_compress_host: bad arguments"); return 1; }
  const int epb = 8 / bits;
  const uint32_t mask = (1u << bits) - 1u;
  const int64_t ob = cols / epb;
  for (int64_t r = 0; r < rows; ++r)
    for (int64_t j = 0; j < ob; ++j) {
      uint32_t b = 0;
      for (int e = 0; e < epb; ++e) b |= (uint32_t(uint8_t(in[r * cols + j * epb + e])) << (bits * e)) & 0xffu;
      (void)mask;
      out[r * ob + j] = int8_t(uint8_t(b));
    }
  return 0;
}

int bb_interleave_host(const int8_t* in, int8_t* out, int64_t nbytes, int bits, int target_bits) {
  if (!in || !out || nbytes % 4 || (bits != 1 && bits != 2 && bits != 4) || (target_bits != 8 && target_bits != 16)) {
    set_error("bb_interleave_host: bad arguments");
    return 1;
  }
  const int layout = layout_from_target(target_bits);
  const int epw = 32 / bits;
  const uint32_t mask = (1u << bits) - 1u;
  const uint32_t* src = reinterpret_cast<const uint32_t*>(in);
  uint32_t* dst = reinterpret_cast<uint32_t*>(out);
  for (int64_t i = 0; i < nbytes / 4; ++i) {
    const uint32_t w = src[i];
    uint32_t o = 0;
    for (int e = 0; e < epw; ++e) o |= ((w >> (bits * e)) & mask) << field_bitpos(e, bits, layout);
    dst[i] = o;
  }
  return 0;
}

int bb_transform_weight_device(const int8_t* w, int8_t* out, int64_t rows, int64_t cols, int bits, int target_bits,
                               void* stream) {
  if (!w || !out || (bits != 1 && bits != 2 && bits != 4) || cols % (32 / bits)) {
    set_error("bb_transform_weight_device: cols must be a multiple of %d", bits ? 32 / bits : 0);
    return 1;
  }
  const int64_t nwords = rows * (cols / (32 / bits));
  const int threads = 256;
  transform_weight_kernel<<<(unsigned)((nwords + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
      w, reinterpret_cast<uint32_t*>(out), rows, cols, bits, layout_from_target(target_bits));
  BB_LAUNCH_CHECK();
  return 0;
}

int bb_retile_weight_device(const int8_t* in, int8_t* out, int64_t rows, int64_t row_bytes, int inverse, void* stream) {
  if (!in || !out || in == out || rows <= 0 || row_bytes <= 0 || rows % BB_TILE_ROWS || row_bytes % BB_TILE_ROW_BYTES ||
      (reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) {
    set_error("bb_retile_weight_device: needs distinct 16-byte aligned buffers, rows %% %d == 0, row_bytes %% %d == 0", BB_TILE_ROWS, BB_TILE_ROW_BYTES);
    return 1;
  }
  const int64_t chunks = rows * (row_bytes / 16);
  const int threads = 256;
  retile_kernel<<<(unsigned)((chunks + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), rows, row_bytes, inverse);
  BB_LAUNCH_CHECK();
  return 0;
}

int bb_repack_gptq_qweight_device(const int32_t* q, int8_t* out, int64_t K, int64_t N, int bits, int target_bits,
                                  void* stream) {
  if (!q || !out || (bits != 2 && bits != 4 && bits != 1) || K % (32 / bits)) { set_error("bb_repack_gptq_qweight_device: bad arguments"); return 1; }
  const int64_t KW = K / (32 / bits);
  dim3 grid((unsigned)((N + 31) / 32), (unsigned)((KW + 31) / 32)), block(32, 8);
  repack_gptq_qweight_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint32_t*>(q),
                                                                        reinterpret_cast<uint32_t*>(out), KW, N, bits,
                                                                        layout_from_target(target_bits));
  BB_LAUNCH_CHECK();
  return 0;
}

int bb_repack_gptq_qzeros_device(const int32_t* qz, const void* scales, void* zeros_out, int64_t groups, int64_t N,
                                 int bits, int zeros_mode, int a_dtype, int v2, void* stream) {
  if (!qz || !zeros_out || (bits != 2 && bits != 4 && bits != 1) || N % (32 / bits)) { set_error("bb_repack_gptq_qzeros_device: bad arguments"); return 1; }
  if (zeros_mode == BB_ZEROS_RESCALE && !scales) { set_error("rescale zeros need scales"); return 1; }
  const int64_t total = zeros_mode == BB_ZEROS_QUANTIZED ? groups * (N * bits / 8) : groups * N;
  const int threads = 256;
  const unsigned blocks = (unsigned)((total + threads - 1) / threads);
  if (a_dtype == BB_BF16)
    repack_gptq_qzeros_kernel<__nv_bfloat16><<<blocks, threads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint32_t*>(qz), (const __nv_bfloat16*)scales, zeros_out, groups, N, bits, zeros_mode, v2);
  else
    repack_gptq_qzeros_kernel<__half><<<blocks, threads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint32_t*>(qz), (const __half*)scales, zeros_out, groups, N, bits, zeros_mode, v2);
  BB_LAUNCH_CHECK();
  return 0;
}

int bb_debug_decode(int kind, int bits, int is_signed, int w_layout, const void* in, void* out, int nwords, void* stream) {
  if ((bits != 2 && bits != 4) || !in || !out) { set_error("bb_debug_decode: bits must be 2 or 4"); return 1; }
  const int zp = is_signed ? (1 << (bits - 1)) : 0;
  const int threads = 128, blocks = (nwords + threads - 1) / threads;
  cudaStream_t s = (cudaStream_t)stream;
  if (kind == 0) debug_decode16_kernel<__half><<<blocks, threads, 0, s>>>(bits, zp, w_layout, (const uint32_t*)in, (__half*)out, nwords);
  else if (kind == 1) debug_decode16_kernel<__nv_bfloat16><<<blocks, threads, 0, s>>>(bits, zp, w_layout, (const uint32_t*)in, (__nv_bfloat16*)out, nwords);
  else if (kind == 2) {
    if (w_layout != BB_LAYOUT_INTERLEAVED_8) { set_error("int8 decode needs the interleaved-8 layout"); return 1; }
    debug_decode8_kernel<<<blocks, threads, 0, s>>>(bits, zp, (const uint32_t*)in, (uint32_t*)out, nwords);
  } else { set_error("bb_debug_decode: bad kind"); return 1; }
  BB_LAUNCH_CHECK();
  return 0;
}

int bb_debug_dequant(int kind, int bits, int is_signed, int w_layout, int mode, const void* in, const void* scale,
                     const void* zeros, const void* qzeros, void* out, int nwords, void* stream) {
  if ((bits != 2 && bits != 4) || !in || !out || mode < 0 || mode > 4 || (kind != 0 && kind != 1)) { set_error("bb_debug_dequant: bad arguments"); return 1; }
  if (w_layout != BB_LAYOUT_COMPRESSED && w_layout != BB_LAYOUT_INTERLEAVED_16) { set_error("bb_debug_dequant: 16-bit layouts only"); return 1; }
  if ((mode >= 1 && !scale) || ((mode == 2 || mode == 3) && !zeros) || (mode == 4 && !qzeros)) { set_error("bb_debug_dequant: missing operand"); return 1; }
  const int zp = is_signed ? (1 << (bits - 1)) : 0;
  const int threads = 128, blocks = (nwords + threads - 1) / threads;
  cudaStream_t s = (cudaStream_t)stream;
#define BB_DQ_GO(TT)                                                                                                              \
  switch (mode) {                                                                                                                 \
    case 0: debug_dequant16_kernel<TT, 0><<<blocks, threads, 0, s>>>(bits, zp, w_layout, (const uint32_t*)in, (const TT*)scale, (const TT*)zeros, (const int*)qzeros, (TT*)out, nwords); break; \
    case 1: debug_dequant16_kernel<TT, 1><<<blocks, threads, 0, s>>>(bits, zp, w_layout, (const uint32_t*)in, (const TT*)scale, (const TT*)zeros, (const int*)qzeros, (TT*)out, nwords); break; \
    case 2: debug_dequant16_kernel<TT, 2><<<blocks, threads, 0, s>>>(bits, zp, w_layout, (const uint32_t*)in, (const TT*)scale, (const TT*)zeros, (const int*)qzeros, (TT*)out, nwords); break; \
    case 3: debug_dequant16_kernel<TT, 3><<<blocks, threads, 0, s>>>(bits, zp, w_layout, (const uint32_t*)in, (const TT*)scale, (const TT*)zeros, (const int*)qzeros, (TT*)out, nwords); break; \
    default: debug_dequant16_kernel<TT, 4><<<blocks, threads, 0, s>>>(bits, zp, w_layout, (const uint32_t*)in, (const TT*)scale, (const TT*)zeros, (const int*)qzeros, (TT*)out, nwords); break; \
  }
  if (kind == 0) { BB_DQ_GO(__half) } else { BB_DQ_GO(__nv_bfloat16) }
#undef BB_DQ_GO
  BB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
