// flpr tcgen05 GEMM / implicit-GEMM convolution for sm_100a.
//
//   D[M,N] = alpha * sum_k A[m,k] * B[n,k]  (+bias, +residual, ReLU)      bf16 operands, fp32 accumulate in TMEM
//
// One 128 x BN output tile per CTA. Warp-specialised:
//   warp 0   : TMA producer  (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 1   : MMA issuer    (one elected thread, tcgen05.mma kind::f16, accumulator in TMEM)
//   warp 2   : TMEM allocator
//   warps 4-7: epilogue      (tcgen05.ld 32x32b -> registers -> fused epilogue -> global)
//
// Operand modes (per operand):
//   OP_KMAJOR  : row-major [rows, K], K contiguous           (forward GEMMs, dgrad A operand)
//   OP_MNMAJOR : row-major [K, rows], rows contiguous        (wgrad operands, dgrad B operand) -> no transposes
//   OP_CONV    : A only. NHWC activation tensor, 4-D TMA box per filter tap; zero padding comes from TMA
//                out-of-bounds fill, so a 3x3 convolution is 9*(C/64) K-blocks of a plain GEMM pipeline.
//
// This is the compute kernel behind ops/gemm.py (classifier, 1x1 / 3x3 bottleneck convs, gallery-vs-query
// similarity). Reference call sites it replaces: models/resnet.py:121-141 (bottleneck convs),
// models/resnet.py:321 (classifier), tools/evaluate.py:100 (per-query torch.mm).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <unordered_map>

#include "ptx.cuh"

namespace flpr {

enum { OP_KMAJOR = 0, OP_MNMAJOR = 1, OP_CONV = 2, OP_TAPFLIP = 3 };
enum { EPI_GENERIC = 0, EPI_LEAN = 1 };

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle atom row
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB

struct GemmParams {
  int M, N, K;          // K = total reduction length (conv: taps * C)
  int kb_per_split;     // K-blocks handled by one blockIdx.z
  int kb_total;
  // epilogue
  void* out;
  long long ldo;
  int out_bf16;
  int trans_out;
  int atomic_add;
  float alpha;
  const float* bias_n;
  const float* bias_m;
  int relu;
  const __nv_bfloat16* residual;
  // conv geometry (A operand in OP_CONV mode)
  int cH, cW, cC, cTH, cNB, cKW, cPadH, cPadW, cTilesPerImg, cSplits;
  long long tap_stride;  // conv wgrad: output column offset per filter tap
  int cTaps;             // KH*KW (OP_TAPFLIP: B column offset = (cTaps-1-tap) * N)
  int cStride;           // convolution stride (A operand: TMA element strides), 1 or 2
  // persistent scheduling
  int tiles_m, tiles_n, tiles_z, tiles_total;
  // fused batch-norm statistics: per (m-tile, epilogue-warp) column partials of sum / sum^2 of the fp32 accumulators
  float* col_part;       // [tiles_m * 4][2][N] or nullptr
  int debug;             // FLPR_GEMM_DEBUG bit mask (bottleneck isolation, results are WRONG when set):
                         //   1 skip the epilogue body, 2 skip global stores, 4 skip TMA loads (MMA on stale smem)
};

template <int BN>
struct SmemLayout {
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  // BN <= 128: ~100 KB per CTA so that TWO CTAs are resident per SM and one CTA's epilogue overlaps the other's
  // MMA main loop (TMEM: 2 x 128 columns). BN = 256 keeps a deeper ring with one CTA per SM.
  static constexpr int STAGES = (BN <= 64) ? 4 : (BN <= 128 ? 3 : 4);
  static constexpr int MIN_CTAS = (BN <= 128) ? 2 : 1;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;  // barriers + tmem ptr + alignment slack
};

// Persistent variant: ring + a dedicated store-staging area (the ring is never idle) + 2 accumulator stages in TMEM.
template <int BN>
struct PersistLayout {
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = (BN <= 128) ? 3 : 4;
  static constexpr int MIN_CTAS = (BN <= 128) ? 2 : 1;
  static constexpr int ACC_STAGES = 2;
  static constexpr int TMEM_COLS = ACC_STAGES * BN;                 // 128/256/512 columns
  // 256-wide tiles (one CTA per SM): 8 epilogue warps, two per TMEM lane quarter, each draining half of the columns
  // - four warps cannot drain a 128 x 256 fp32 tile within the 4096 MMA cycles of a K = 512 tile.
  static constexpr int EPI_WARPS = (BN == 256) ? 8 : 4;
  static constexpr int THREADS = 128 + 32 * EPI_WARPS;
  static constexpr int STORE_OFFSET = STAGES * STAGE_BYTES;         // epilogue warps x 32 rows x 80 B
  static constexpr int STORE_BYTES = EPI_WARPS * 32 * 80;
  static constexpr int BAR_OFFSET = STORE_OFFSET + STORE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;
};

// ------------------------------------------------------------------------------------------------- shared pieces
// TMA loads of one K-block (BK = 64) of both operands into one ring stage.
template <int BN, int A_MODE, int B_MODE, int CG = 1>
__device__ __forceinline__ void load_kblock(const GemmParams& p, const CUtensorMap* tmA, const CUtensorMap* tmB,
                                            uint8_t* sa, uint8_t* sb, uint64_t* bar, int kb, int mt, int m0, int n0,
                                            int ztap) {
  // CG == 2: CTA-pair kernel, `bar` is this CTA's copy of the full barrier; the bytes are credited to the leader's
  auto tma_load_2d = [](void* d, const CUtensorMap* m, uint64_t* b, int c0, int c1) {
    if constexpr (CG == 2) flpr::tma2_load_2d(d, m, b, c0, c1); else flpr::tma_load_2d(d, m, b, c0, c1);
  };
  auto tma_load_4d = [](void* d, const CUtensorMap* m, uint64_t* b, int c0, int c1, int c2, int c3) {
    if constexpr (CG == 2) flpr::tma2_load_4d(d, m, b, c0, c1, c2, c3); else flpr::tma_load_4d(d, m, b, c0, c1, c2, c3);
  };
  int tapA = 0;
  if constexpr (A_MODE == OP_KMAJOR) {
    tma_load_2d(sa, tmA, bar, kb * BK, m0);
  } else if constexpr (A_MODE == OP_MNMAJOR) {
    tma_load_2d(sa, tmA, bar, m0, kb * BK);
    tma_load_2d(sa + 64 * 128, tmA, bar, m0 + 64, kb * BK);
  } else {
    const int chunks = p.cC / BK;
    tapA = kb / chunks;
    const int cc = kb - tapA * chunks;
    const int kh = tapA / p.cKW;
    const int kw = tapA - kh * p.cKW;
    int img0, h0;
    if (p.cTilesPerImg <= 1) {
      img0 = mt * p.cNB;
      h0 = 0;
    } else {
      img0 = mt / p.cTilesPerImg;
      h0 = (mt - img0 * p.cTilesPerImg) * p.cTH;
    }
    tma_load_4d(sa, tmA, bar, cc * BK, kw - p.cPadW, h0 * p.cStride + kh - p.cPadH, img0);
  }
  if constexpr (B_MODE == OP_KMAJOR) {
    tma_load_2d(sb, tmB, bar, kb * BK, n0);
  } else if constexpr (B_MODE == OP_MNMAJOR) {
#pragma unroll
    for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * 64 * 128, tmB, bar, n0 + j * 64, kb * BK);
  } else if constexpr (B_MODE == OP_TAPFLIP) {
    // data-gradient of a convolution straight from the forward weight W[Cout, taps, Cin]: for A tap t the B operand
    // is W[:, taps-1-t, :] viewed as [K = Cout rows (stride taps*Cin), N = Cin contiguous] -> no flipped copy.
    const int chunks = p.cC / BK;                         // A channels (= Cout of the forward conv) per tap
    const int cc = kb - tapA * chunks;
    const int col0 = (p.cTaps - 1 - tapA) * p.N + n0;
#pragma unroll
    for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * 64 * 128, tmB, bar, col0 + j * 64, cc * BK);
  } else {
    // weight-gradient of a convolution: K runs over pixels (64 per block), the filter tap (ztap) shifts the 4-D
    // box, zero padding comes from TMA out-of-bounds fill
    const int kh = ztap / p.cKW;
    const int kw = ztap - kh * p.cKW;
    const int p0 = kb * BK;
    const int hw = p.cH * p.cW;
    const int img = p0 / hw;
    const int h0 = (p0 - img * hw) / p.cW;
#pragma unroll
    for (int j = 0; j < BN / 64; ++j)
      tma_load_4d(sb + j * 64 * 128, tmB, bar, n0 + j * 64, kw - p.cPadW, h0 + kh - p.cPadH, img);
  }
}

// The four UMMA_K = 16 steps of one K-block. Called by ONE thread.
template <int BN, int A_MODE, int B_MODE, int CG = 1>
__device__ __forceinline__ void mma_kblock(uint32_t sa, uint32_t tmem_acc, bool first_kb) {
  // CG == 2: BN is the pair's N (each CTA holds BN/2 rows of B and 128 rows of A); the instruction covers M = 256
  constexpr uint32_t idesc = make_idesc_bf16(BM * CG, BN, A_MODE == OP_MNMAJOR, B_MODE != OP_KMAJOR);
  const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
  for (int k = 0; k < BK / UMMA_K; ++k) {
    uint64_t da, db;
    if constexpr (A_MODE == OP_MNMAJOR)
      da = make_smem_desc_sw128(sa + k * (UMMA_K * 128), 64 * 128, 1024);
    else
      da = make_smem_desc_sw128(sa + k * (UMMA_K * 2), 16, 1024);
    if constexpr (B_MODE != OP_KMAJOR)
      db = make_smem_desc_sw128(sb + k * (UMMA_K * 128), 64 * 128, 1024);
    else
      db = make_smem_desc_sw128(sb + k * (UMMA_K * 2), 16, 1024);
    if constexpr (CG == 2)
      umma2_f16(tmem_acc, da, db, idesc, (!first_kb || k != 0) ? 1u : 0u);
    else
      umma_f16(tmem_acc, da, db, idesc, (!first_kb || k != 0) ? 1u : 0u);
  }
}

// Column sums over the 32 rows held by a warp (lane = row, v[j] = column j) with a butterfly of 31 shuffles:
// afterwards v[0] on lane l is the total of column l.
__device__ __forceinline__ void warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int j = 0; j < s; ++j) {
      const float send = up ? v[j] : v[j + s];
      const float keep = up ? v[j + s] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
}

// Epilogue of columns [cb, ce) of one 128 x BN tile by one warp (q = TMEM lane quarter): TMEM -> registers ->
// alpha / bias / residual / ReLU -> global (bf16 / fp32 / transposed / atomic). `have_acc` false = zero accumulator.
// The TMEM load of chunk i+1 is issued before chunk i is processed (two register buffers).
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, const uint32_t (&r)[32], int mt, int m0, int n0,
                                               int c0, long long tap_off, int q, int lane, uint8_t* stage_buf,
                                               float bm, bool use_stage) {
  const int row = m0 + q * 32 + lane;
  if (n0 + c0 >= p.N) return;
  if (p.debug & 1) return;
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    float x = __uint_as_float(r[j]) * p.alpha + bm;
    const int col = n0 + c0 + j;
    if (p.bias_n != nullptr && col < p.N) x += p.bias_n[col];
    v[j] = x;
  }
  const bool stats_from_stage = p.col_part != nullptr && use_stage && (n0 + c0 + 32 <= p.N);
  if (p.col_part != nullptr && !stats_from_stage) {
    // fused batch-norm statistics of the raw fp32 conv output (rows >= M are exact zeros: TMA OOB fill)
    // (of the value that is stored: bf16-rounded when the output is bf16, like the staged path below)
    float s1[32], s2[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float x = p.out_bf16 ? __bfloat162float(__float2bfloat16(v[j])) : v[j];
      s1[j] = x;
      s2[j] = x * x;
    }
    warp_colsum32(s1, lane);
    warp_colsum32(s2, lane);
    const int col = n0 + c0 + lane;
    if (col < p.N) {
      float* dst = p.col_part + ((size_t)(mt * 4 + q) * 2) * p.N + col;
      dst[0] = s1[0];
      dst[p.N] = s2[0];
    }
  }
  if (p.trans_out) {
    // out[col * ldo + row]: lanes are contiguous in memory -> coalesced per column
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int col = n0 + c0 + j;
      if (row < p.M && col < p.N) {
        const long long idx = (long long)col * p.ldo + row;
        float x = v[j];
        if (p.residual != nullptr) x += __bfloat162float(p.residual[idx]);
        if (p.relu) x = fmaxf(x, 0.f);
        if (p.atomic_add)
          atomicAdd(reinterpret_cast<float*>(p.out) + idx, x);
        else if (p.out_bf16)
          reinterpret_cast<__nv_bfloat16*>(p.out)[idx] = __float2bfloat16(x);
        else
          reinterpret_cast<float*>(p.out)[idx] = x;
      }
    }
  } else if (row < p.M || use_stage) {
    const long long base = (long long)row * p.ldo + n0 + c0 + tap_off;
    const bool full = (n0 + c0 + 32 <= p.N);
    if (p.residual != nullptr && row < p.M) {
      if (full && ((base & 7) == 0)) {
        const uint4* rp = reinterpret_cast<const uint4*>(p.residual + base);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          uint4 u = rp[j4];
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            float2 f = __bfloat1622float2(h[t]);
            v[j4 * 8 + t * 2] += f.x;
            v[j4 * 8 + t * 2 + 1] += f.y;
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (n0 + c0 + j < p.N) v[j] += __bfloat162float(p.residual[base + j]);
      }
    }
    if (p.relu) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    if (p.atomic_add) {
      float* o = reinterpret_cast<float*>(p.out) + base;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n0 + c0 + j < p.N) atomicAdd(o + j, v[j]);
    } else if (p.out_bf16) {
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + base;
      if (full && ((base & 7) == 0)) {
        uint4 pk[4];
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk[j4]);
#pragma unroll
          for (int t = 0; t < 4; ++t) h[t] = __floats2bfloat162_rn(v[j4 * 8 + t * 2], v[j4 * 8 + t * 2 + 1]);
        }
        if (use_stage) {
          // transpose through this warp's private staging tile so that 4 lanes write one row's 64 B contiguously
          const uint32_t sbase = smem_u32(stage_buf);
          const uint32_t st = sbase + lane * 80;
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) sts_128(st + j4 * 16, pk[j4]);
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int rr = 8 * j + (lane >> 2);
            const uint4 u = lds_128(sbase + rr * 80 + (lane & 3) * 16);
            const int grow = m0 + q * 32 + rr;
            if (grow < p.M && !(p.debug & 2))
              *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + (long long)grow * p.ldo + n0 +
                                        c0 + tap_off + (lane & 3) * 8) = u;
          }
          if (stats_from_stage) {
            // fused batch-norm statistics: lane = column, walk the 32 staged rows (the bf16 values that are stored;
            // rows >= M hold exact zeros). 32 conflict-free 2-byte smem loads instead of 62 shuffles.
            float a1 = 0.f, a2 = 0.f, b1 = 0.f, b2 = 0.f;
#pragma unroll
            for (int rr = 0; rr < 32; rr += 2) {
              const float x = lds_bf16(sbase + rr * 80 + lane * 2);
              const float y = lds_bf16(sbase + (rr + 1) * 80 + lane * 2);
              a1 += x;
              a2 = fmaf(x, x, a2);
              b1 += y;
              b2 = fmaf(y, y, b2);
            }
            float* dst = p.col_part + ((size_t)(mt * 4 + q) * 2) * p.N + n0 + c0 + lane;
            dst[0] = a1 + b1;
            dst[p.N] = a2 + b2;
          }
          __syncwarp();
        } else {
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) reinterpret_cast<uint4*>(o)[j4] = pk[j4];
        }
      } else if (row < p.M) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (n0 + c0 + j < p.N) o[j] = __float2bfloat16(v[j]);
      }
    } else {
      float* o = reinterpret_cast<float*>(p.out) + base;
      if (full && ((base & 3) == 0)) {
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4)
          reinterpret_cast<float4*>(o)[j4] = make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (n0 + c0 + j < p.N) o[j] = v[j];
      }
    }
  }
}

template <bool PIPE>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, uint32_t tmem_acc, bool have_acc, int mt, int m0,
                                              int n0, int cb, int ce, long long tap_off, int q, int lane,
                                              uint8_t* stage_buf) {
  const int row = m0 + q * 32 + lane;
  const float bm = (p.bias_m != nullptr && row < p.M) ? p.bias_m[row] : 0.f;
  // bf16 row-major output without residual: stores are staged through smem so that each store instruction writes
  // 8 full 64-byte row segments
  const bool use_stage = p.out_bf16 && !p.trans_out && !p.atomic_add && p.residual == nullptr &&
                         ((p.ldo & 7) == 0) && (((n0 + tap_off) & 7) == 0) &&
                         ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
  const uint32_t tbase = tmem_acc + (uint32_t(q * 32) << 16);
  uint32_t ra[32], rb[32];
  if (!have_acc) {
#pragma unroll
    for (int j = 0; j < 32; ++j) ra[j] = 0u;
    for (int c0 = cb; c0 < ce; c0 += 32) epilogue_chunk(p, ra, mt, m0, n0, c0, tap_off, q, lane, stage_buf, bm, use_stage);
    return;
  }
  if constexpr (!PIPE) {       // 128-register budget (two CTAs per SM): one register buffer
#pragma unroll 1
    for (int c0 = cb; c0 < ce; c0 += 32) {
      tmem_ld_32x32b_x32(tbase + uint32_t(c0), ra);
      tmem_ld_wait();
      epilogue_chunk(p, ra, mt, m0, n0, c0, tap_off, q, lane, stage_buf, bm, use_stage);
    }
    return;
  }
  tmem_ld_32x32b_x32(tbase + uint32_t(cb), ra);
#pragma unroll 1
  for (int c0 = cb; c0 < ce; c0 += 64) {
    tmem_ld_wait();
    if (c0 + 32 < ce) tmem_ld_32x32b_x32(tbase + uint32_t(c0 + 32), rb);
    epilogue_chunk(p, ra, mt, m0, n0, c0, tap_off, q, lane, stage_buf, bm, use_stage);
    if (c0 + 32 < ce) {
      tmem_ld_wait();
      if (c0 + 64 < ce) tmem_ld_32x32b_x32(tbase + uint32_t(c0 + 64), ra);
      epilogue_chunk(p, rb, mt, m0, n0, c0 + 32, tap_off, q, lane, stage_buf, bm, use_stage);
    }
  }
}

// Lean epilogue (EPI_LEAN): the three hot output kinds - bf16 store (+ fused BN statistics), fp32 store, fp32
// atomic accumulate (split-K) - all row-major, no bias / residual / ReLU / transpose, N a multiple of 32. Every chunk
// goes through the warp's 32 x 64-byte staging tile so that one store instruction covers 8 full 64-byte row segments
// (a lane-per-row store would touch 32 different lines per instruction; the atomic path uses red.v4.f32).
// A separate, small code path: the generic epilogue above inlines ~7k instructions, which thrashes the instruction
// cache when it sits in the inner loop of the persistent kernel.
template <bool PIPE>
__device__ __forceinline__ void epilogue_tile_lean(const GemmParams& p, uint32_t tmem_acc, int mt, int m0, int n0,
                                                   int cb, int ce, long long tap_off, int q, int lane,
                                                   uint32_t sbase) {
  const uint32_t tbase = tmem_acc + (uint32_t(q * 32) << 16);
  const uint32_t st = sbase + lane * 80;
  const int sub_row = lane >> 2, sub_col = lane & 3;
  const float alpha = p.alpha;
  uint32_t ra[32], rb[32];
  auto chunk = [&](const uint32_t (&r)[32], int c0) {
    if ((p.debug & 1) || n0 + c0 >= p.N) return;       // (tile wider than the matrix: nothing to store)
    if (p.out_bf16) {
      uint4 pk[4];
      if (p.bias_n == nullptr && p.residual == nullptr && !p.relu) {
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk[j4]);
#pragma unroll
          for (int t = 0; t < 4; ++t)
            h[t] = __floats2bfloat162_rn(__uint_as_float(r[j4 * 8 + t * 2]) * alpha,
                                         __uint_as_float(r[j4 * 8 + t * 2 + 1]) * alpha);
        }
      } else {
        // inference epilogue of the BN-folded trunk: + bias[col] (+ residual[row, col]) -> ReLU
        const int grow_l = m0 + q * 32 + lane;
        const bool has_res = p.residual != nullptr && grow_l < p.M;
        const uint4* rp = reinterpret_cast<const uint4*>(p.residual + (long long)grow_l * p.ldo + n0 + c0 + tap_off);
        const float4* bp = reinterpret_cast<const float4*>(p.bias_n + n0 + c0);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          float v[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) v[t] = __uint_as_float(r[j4 * 8 + t]) * alpha;
          if (p.bias_n != nullptr) {
            const float4 b0 = bp[j4 * 2], b1 = bp[j4 * 2 + 1];
            v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
            v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
          }
          if (has_res) {
            const uint4 u = rp[j4];
            const __nv_bfloat162* rh = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float2 f = __bfloat1622float2(rh[t]);
              v[2 * t] += f.x;
              v[2 * t + 1] += f.y;
            }
          }
          if (p.relu) {
#pragma unroll
            for (int t = 0; t < 8; ++t) v[t] = fmaxf(v[t], 0.f);
          }
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk[j4]);
#pragma unroll
          for (int t = 0; t < 4; ++t) h[t] = __floats2bfloat162_rn(v[2 * t], v[2 * t + 1]);
        }
      }
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4) sts_128(st + j4 * 16, pk[j4]);
      __syncwarp();
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + n0 + c0 + tap_off + sub_col * 8;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int rr = 8 * j + sub_row;
        const uint4 u = lds_128(sbase + rr * 80 + sub_col * 16);
        const int grow = m0 + q * 32 + rr;
        if (grow < p.M) *reinterpret_cast<uint4*>(o + (long long)grow * p.ldo) = u;
      }
      if (p.col_part != nullptr) {
        // fused batch-norm statistics of the stored (bf16) values: lane = column, 32 staged rows (rows >= M are 0)
        float a1 = 0.f, a2 = 0.f, b1 = 0.f, b2 = 0.f;
#pragma unroll
        for (int rr = 0; rr < 32; rr += 2) {
          const float x = lds_bf16(sbase + rr * 80 + lane * 2);
          const float y = lds_bf16(sbase + (rr + 1) * 80 + lane * 2);
          a1 += x;
          a2 = fmaf(x, x, a2);
          b1 += y;
          b2 = fmaf(y, y, b2);
        }
        float* dst = p.col_part + ((size_t)(mt * 4 + q) * 2) * p.N + n0 + c0 + lane;
        dst[0] = a1 + b1;
        dst[p.N] = a2 + b2;
      }
      __syncwarp();
    } else {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          uint4 u;
          u.x = __float_as_uint(__uint_as_float(r[half * 16 + j4 * 4 + 0]) * alpha);
          u.y = __float_as_uint(__uint_as_float(r[half * 16 + j4 * 4 + 1]) * alpha);
          u.z = __float_as_uint(__uint_as_float(r[half * 16 + j4 * 4 + 2]) * alpha);
          u.w = __float_as_uint(__uint_as_float(r[half * 16 + j4 * 4 + 3]) * alpha);
          sts_128(st + j4 * 16, u);
        }
        __syncwarp();
        float* o = reinterpret_cast<float*>(p.out) + n0 + c0 + half * 16 + tap_off + sub_col * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int rr = 8 * j + sub_row;
          const uint4 u = lds_128(sbase + rr * 80 + sub_col * 16);
          const int grow = m0 + q * 32 + rr;
          if (grow < p.M) {
            float* dst = o + (long long)grow * p.ldo;
            if (p.atomic_add)
              red_add_v4_f32(dst, u);
            else
              *reinterpret_cast<uint4*>(dst) = u;
          }
        }
        __syncwarp();
      }
    }
  };
  if constexpr (!PIPE) {
#pragma unroll 1
    for (int c0 = cb; c0 < ce; c0 += 32) {
      tmem_ld_32x32b_x32(tbase + uint32_t(c0), ra);
      tmem_ld_wait();
      chunk(ra, c0);
    }
  } else {
    tmem_ld_32x32b_x32(tbase + uint32_t(cb), ra);
#pragma unroll 1
    for (int c0 = cb; c0 < ce; c0 += 64) {
      tmem_ld_wait();
      if (c0 + 32 < ce) tmem_ld_32x32b_x32(tbase + uint32_t(c0 + 32), rb);
      chunk(ra, c0);
      if (c0 + 32 < ce) {
        tmem_ld_wait();
        if (c0 + 64 < ce) tmem_ld_32x32b_x32(tbase + uint32_t(c0 + 64), ra);
        chunk(rb, c0 + 32);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------- classic kernel
// One tile per CTA (grid = tiles_n x tiles_m x splits). Kept as the fallback / comparison path
// (FLPR_GEMM_PERSIST=0) of the persistent kernel below.
template <int BN, int A_MODE, int B_MODE>
__global__ void __launch_bounds__(256, SmemLayout<BN>::MIN_CTAS)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const GemmParams p) {
  using L = SmemLayout<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + L::STAGES;
  uint64_t* tmem_full_bar = empty_bar + L::STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp_id = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * BM;
  const int zsplit = (B_MODE == OP_CONV) ? (blockIdx.z % p.cSplits) : blockIdx.z;
  const int ztap = (B_MODE == OP_CONV) ? (blockIdx.z / p.cSplits) : 0;
  const int kb_begin = zsplit * p.kb_per_split;
  const int kb_end = min(kb_begin + p.kb_per_split, p.kb_total);
  const int num_kb = kb_end - kb_begin;

  if (warp_id == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp_id == 1 && lane == 0) {
    for (int s = 0; s < L::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp_id == 2) {
    tmem_alloc(tmem_ptr_smem, BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp_id == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < num_kb; ++i) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * L::STAGE_BYTES;
        mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
        load_kblock<BN, A_MODE, B_MODE>(p, &tmA, &tmB, sa, sa + A_STAGE_BYTES, &full_bar[stage], kb_begin + i,
                                        blockIdx.y, m0, n0, ztap);
        if (++stage == L::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp_id == 1) {
    // ===================== MMA issuer =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < num_kb; ++i) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) mma_kblock<BN, A_MODE, B_MODE>(smem_u32(smem + stage * L::STAGE_BYTES), tmem_base, i == 0);
      __syncwarp();
      if (elect_one()) umma_commit(&empty_bar[stage]);
      __syncwarp();
      if (++stage == L::STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
    if (elect_one()) umma_commit(tmem_full_bar);
    __syncwarp();
  } else if (warp_id >= 4) {
    // ===================== epilogue =====================
    const int q = warp_id & 3;  // TMEM lane quarter this warp may access
    if (num_kb > 0) {
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
    }
    const long long tap_off = (B_MODE == OP_CONV) ? (long long)ztap * p.tap_stride : 0;
    // pipeline stage 0 is free once the accumulator barrier has fired -> store staging
    epilogue_tile<false>(p, tmem_base, num_kb > 0, blockIdx.y, m0, n0, 0, BN, tap_off, q, lane, smem + q * (32 * 80));
  }

  tc_fence_before();
  __syncthreads();
  if (warp_id == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

// ------------------------------------------------------------------------------------------------- persistent kernel
// grid = min(#tiles, SMs x resident CTAs); every CTA walks tiles t = blockIdx.x, + gridDim.x, ... (n fastest, so
// that concurrently running CTAs share the A rows in L2). Two accumulator stages in TMEM: the epilogue warps drain
// tile i (TMEM -> registers -> global) while the MMA warp already accumulates tile i+1, and the TMA producer runs
// ahead across tile boundaries, so neither the pipeline fill nor the epilogue is exposed for short-K GEMMs.
template <int BN, int A_MODE, int B_MODE, int EPI>
__global__ void __launch_bounds__(PersistLayout<BN>::THREADS, PersistLayout<BN>::MIN_CTAS)
gemm_bf16_tcgen05_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                    const GemmParams p) {
  using L = PersistLayout<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + L::STAGES;
  uint64_t* tmem_full_bar = empty_bar + L::STAGES;          // [ACC_STAGES]
  uint64_t* tmem_empty_bar = tmem_full_bar + L::ACC_STAGES;  // [ACC_STAGES]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + L::ACC_STAGES);

  const int warp_id = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp_id == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp_id == 1 && lane == 0) {
    for (int s = 0; s < L::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < L::ACC_STAGES; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], L::EPI_WARPS);   // one arrival per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp_id == 2) {
    tmem_alloc(tmem_ptr_smem, L::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int per_z = p.tiles_m * p.tiles_n;

  if (warp_id == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.tiles_total; t += gridDim.x) {
        const int z = t / per_z;
        const int r = t - z * per_z;
        const int mt = r / p.tiles_n;
        const int nt = r - mt * p.tiles_n;
        const int zsplit = (B_MODE == OP_CONV) ? (z % p.cSplits) : z;
        const int ztap = (B_MODE == OP_CONV) ? (z / p.cSplits) : 0;
        const int kb_begin = zsplit * p.kb_per_split;
        const int kb_end = min(kb_begin + p.kb_per_split, p.kb_total);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::STAGE_BYTES;
          if (p.debug & 4) {
            mbar_arrive(&full_bar[stage]);
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
            load_kblock<BN, A_MODE, B_MODE>(p, &tmA, &tmB, sa, sa + A_STAGE_BYTES, &full_bar[stage], kb, mt, mt * BM,
                                            nt * BN, ztap);
          }
          if (++stage == L::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp_id == 1) {
    // ===================== MMA issuer =====================
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < p.tiles_total; t += gridDim.x) {
      const int z = t / per_z;
      const int zsplit = (B_MODE == OP_CONV) ? (z % p.cSplits) : z;
      const int kb_begin = zsplit * p.kb_per_split;
      const int num_kb = min(kb_begin + p.kb_per_split, p.kb_total) - kb_begin;
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);       // epilogue has drained this accumulator stage
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + uint32_t(acc * BN);
      for (int i = 0; i < num_kb; ++i) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) mma_kblock<BN, A_MODE, B_MODE>(smem_u32(smem + stage * L::STAGE_BYTES), tmem_acc, i == 0);
        __syncwarp();
        if (elect_one()) umma_commit(&empty_bar[stage]);
        __syncwarp();
        if (++stage == L::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (elect_one()) umma_commit(&tmem_full_bar[acc]);
      __syncwarp();
      if (++acc == L::ACC_STAGES) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else if (warp_id >= 4) {
    // ===================== epilogue =====================
    const int q = warp_id & 3;
    const int part = (warp_id - 4) >> 2;                      // which column slice of the tile this warp drains
    constexpr int COLS_PER_WARP = BN / (L::EPI_WARPS / 4);
    uint8_t* stage_buf = smem + L::STORE_OFFSET + (warp_id - 4) * (32 * 80);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < p.tiles_total; t += gridDim.x) {
      const int z = t / per_z;
      const int r = t - z * per_z;
      const int mt = r / p.tiles_n;
      const int nt = r - mt * p.tiles_n;
      const int ztap = (B_MODE == OP_CONV) ? (z / p.cSplits) : 0;
      const long long tap_off = (B_MODE == OP_CONV) ? (long long)ztap * p.tap_stride : 0;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      if constexpr (EPI == EPI_LEAN)
        epilogue_tile_lean<(BN == 256)>(p, tmem_base + uint32_t(acc * BN), mt, mt * BM, nt * BN, part * COLS_PER_WARP,
                                        (part + 1) * COLS_PER_WARP, tap_off, q, lane, smem_u32(stage_buf));
      else
        epilogue_tile<(BN == 256)>(p, tmem_base + uint32_t(acc * BN), true, mt, mt * BM, nt * BN,
                                   part * COLS_PER_WARP, (part + 1) * COLS_PER_WARP, tap_off, q, lane, stage_buf);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      if (++acc == L::ACC_STAGES) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_id == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, L::TMEM_COLS);
  }
}


// ------------------------------------------------------------------------------------------------- CTA-pair kernel
// cta_group::2: two CTAs of a cluster (one TPC) work on ONE 256 x 256 tile. Each CTA stages 128 rows of A (its half of
// M) and 128 rows of B (its half of N) - 32 KB per K-block instead of 48 KB for a 128 x 256 single-CTA tile, i.e. a
// third less L2 -> smem traffic per FLOP, and a 6-stage ring - the leader CTA issues tcgen05.mma.cta_group::2 over both
// CTAs' shared memory, and each CTA's TMEM receives the accumulators of its 128 rows x 256 columns. Barriers:
//   full[s]   (leader)  : 2 arrivals (each CTA's producer) + the TMA bytes of both CTAs
//   empty[s]  (per CTA) : tcgen05.commit multicast from the leader's MMA thread
//   tmem_full (per CTA) : tcgen05.commit multicast;   tmem_empty (leader): one arrival per epilogue warp of BOTH CTAs
// Persistent over pairs; lean epilogue only (the host falls back to the single-CTA kernel otherwise).
struct PairLayout {
  static constexpr int BN = 256;                       // N of the pair tile
  static constexpr int BNH = 128;                      // B rows staged per CTA
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + BNH * BK * 2;   // 32 KB
  static constexpr int STAGES = 6;
  static constexpr int ACC_STAGES = 2;
  static constexpr int TMEM_COLS = 512;
  static constexpr int EPI_WARPS = 8;
  static constexpr int THREADS = 128 + 32 * EPI_WARPS;
  static constexpr int STORE_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int STORE_BYTES = EPI_WARPS * 32 * 80;
  static constexpr int BAR_OFFSET = STORE_OFFSET + STORE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;
};

template <int A_MODE, int B_MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PairLayout::THREADS, 1)
gemm_bf16_tcgen05_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                              const GemmParams p) {
  using L = PairLayout;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + L::STAGES;
  uint64_t* tmem_full_bar = empty_bar + L::STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + L::ACC_STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + L::ACC_STAGES);

  const int warp_id = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;

  if (warp_id == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp_id == 1 && lane == 0) {
    for (int s = 0; s < L::STAGES; ++s) {
      mbar_init(&full_bar[s], 2);                       // the two producers (leader: + expect_tx of both CTAs)
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < L::ACC_STAGES; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 2 * L::EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp_id == 2) {
    tmem_alloc2(tmem_ptr_smem, L::TMEM_COLS);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                   // barrier inits of both CTAs visible before any remote access
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int tiles_m2 = (p.tiles_m + 1) >> 1;            // p.tiles_m counts 128-row tiles
  const int per_z = tiles_m2 * p.tiles_n;
  const int tiles_total = per_z * p.tiles_z;

  if (warp_id == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = pair; t < tiles_total; t += npairs) {
        const int z = t / per_z;
        const int r = t - z * per_z;
        const int mt2 = r / p.tiles_n;
        const int nt = r - mt2 * p.tiles_n;
        const int mt = mt2 * 2 + (int)rank;
        const int zsplit = (B_MODE == OP_CONV) ? (z % p.cSplits) : z;
        const int ztap = (B_MODE == OP_CONV) ? (z / p.cSplits) : 0;
        const int kb_begin = zsplit * p.kb_per_split;
        const int kb_end = min(kb_begin + p.kb_per_split, p.kb_total);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::STAGE_BYTES;
          if (leader)
            mbar_arrive_expect_tx(&full_bar[stage], 2 * L::STAGE_BYTES);
          load_kblock<L::BNH, A_MODE, B_MODE, 2>(p, &tmA, &tmB, sa, sa + A_STAGE_BYTES, &full_bar[stage], kb, mt,
                                                 mt * BM, nt * L::BN + (int)rank * L::BNH, ztap);
          if (!leader) mbar_arrive_cluster(&full_bar[stage], 0);
          if (++stage == L::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp_id == 1 && leader) {
    // ===================== MMA issuer (leader CTA only) =====================
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = pair; t < tiles_total; t += npairs) {
      const int z = t / per_z;
      const int zsplit = (B_MODE == OP_CONV) ? (z % p.cSplits) : z;
      const int kb_begin = zsplit * p.kb_per_split;
      const int num_kb = min(kb_begin + p.kb_per_split, p.kb_total) - kb_begin;
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + uint32_t(acc * L::BN);
      for (int i = 0; i < num_kb; ++i) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) mma_kblock<L::BN, A_MODE, B_MODE, 2>(smem_u32(smem + stage * L::STAGE_BYTES), tmem_acc, i == 0);
        __syncwarp();
        if (elect_one()) umma2_commit_multicast(&empty_bar[stage]);
        __syncwarp();
        if (++stage == L::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (elect_one()) umma2_commit_multicast(&tmem_full_bar[acc]);
      __syncwarp();
      if (++acc == L::ACC_STAGES) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else if (warp_id >= 4) {
    // ===================== epilogue (both CTAs, own 128 rows x 256 columns) =====================
    const int q = warp_id & 3;
    const int part = (warp_id - 4) >> 2;
    constexpr int COLS_PER_WARP = L::BN / (L::EPI_WARPS / 4);
    const uint32_t sbase = smem_u32(smem + L::STORE_OFFSET + (warp_id - 4) * (32 * 80));
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = pair; t < tiles_total; t += npairs) {
      const int z = t / per_z;
      const int r = t - z * per_z;
      const int mt2 = r / p.tiles_n;
      const int nt = r - mt2 * p.tiles_n;
      const int mt = mt2 * 2 + (int)rank;
      const int ztap = (B_MODE == OP_CONV) ? (z / p.cSplits) : 0;
      const long long tap_off = (B_MODE == OP_CONV) ? (long long)ztap * p.tap_stride : 0;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      if (mt < p.tiles_m)
        epilogue_tile_lean<true>(p, tmem_base + uint32_t(acc * L::BN), mt, mt * BM, nt * L::BN, part * COLS_PER_WARP,
                                 (part + 1) * COLS_PER_WARP, tap_off, q, lane, sbase);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tmem_empty_bar[acc], 0);
      if (++acc == L::ACC_STAGES) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                   // nobody may still target the peer's smem / TMEM
  if (warp_id == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, L::TMEM_COLS);
  }
}

// =============================================================================== host side
static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
static std::mutex g_mu;
static char g_err[512] = {0};

static int set_err(const char* what, int code) {
  snprintf(g_err, sizeof(g_err), "%s (code %d)", what, code);
  return code ? code : -1;
}

static bool load_encode() {
  if (g_encode) return true;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr) return false;
  g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  return true;
}

struct MapKey {
  const void* ptr;
  uint64_t d0, d1, d2, d3, s1, s2, s3;
  uint32_t b0, b1, b2, b3, rank, estride, pad_;
  bool operator==(const MapKey& o) const { return memcmp(this, &o, sizeof(MapKey)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    size_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(MapKey) / 8; ++i) h = (h ^ w[i]) * 1099511628211ull;
    return h;
  }
};
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;

// rank-2 or rank-4 bf16 tensor map with 128B swizzle, zero OOB fill.
static int get_map(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, uint32_t spatial_stride = 1) {
  MapKey k;
  memset(&k, 0, sizeof(k));
  k.ptr = ptr;
  k.rank = rank;
  k.estride = spatial_stride;
  k.d0 = dims[0];
  k.d1 = dims[1];
  k.b0 = box[0];
  k.b1 = box[1];
  k.s1 = strides_bytes[0];
  if (rank == 4) {
    k.d2 = dims[2];
    k.d3 = dims[3];
    k.b2 = box[2];
    k.b3 = box[3];
    k.s2 = strides_bytes[1];
    k.s3 = strides_bytes[2];
  }
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_maps.find(k);
  if (it != g_maps.end()) {
    *out = it->second;
    return 0;
  }
  if (!load_encode()) return set_err("cuTensorMapEncodeTiled entry point unavailable", -2);
  cuuint64_t gdim[4];
  cuuint64_t gstr[3];
  cuuint32_t bx[4];
  cuuint32_t es[4] = {1, 1, 1, 1};
  if (rank == 4 && spatial_stride > 1) es[1] = es[2] = spatial_stride;   // strided convolution: every s-th pixel
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
  }
  for (int i = 0; i < rank - 1; ++i) gstr[i] = strides_bytes[i];
  CUtensorMap m;
  CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), gdim, gstr, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err("cuTensorMapEncodeTiled failed", (int)r);
  if (g_maps.size() > 4096) g_maps.clear();
  g_maps.emplace(k, m);
  *out = m;
  return 0;
}

static int g_persist = -1;   // FLPR_GEMM_PERSIST (default 1): persistent kernel with overlapped epilogue
static int g_sms = 0;

static bool use_persist() {
  if (g_persist < 0) {
    const char* e = getenv("FLPR_GEMM_PERSIST");
    g_persist = (e == nullptr || e[0] != '0') ? 1 : 0;
  }
  return g_persist == 1;
}

static int sm_count() {
  if (g_sms == 0) {
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    g_sms = n;
  }
  return g_sms;
}

static int g_pair = -1;      // FLPR_GEMM_2CTA (default 1): cta_group::2 kernel on the shapes it covers

static bool use_pair() {
  if (g_pair < 0) {
    const char* e = getenv("FLPR_GEMM_2CTA");
    g_pair = (e == nullptr || e[0] != '0') ? 1 : 0;
  }
  return g_pair == 1 && use_persist();
}

// FLPR_GEMM_GENERIC_EPI / FLPR_GEMM_DEBUG are read ONCE (no getenv on the launch path: launches come from several
// client threads, and getenv is not safe against a concurrent setenv elsewhere in the process).
static int g_generic_epi = -1;
static int g_debug = -1;

static bool force_generic_epi() {
  if (g_generic_epi < 0) g_generic_epi = getenv("FLPR_GEMM_GENERIC_EPI") != nullptr ? 1 : 0;
  return g_generic_epi == 1;
}

static int debug_flags() {
  if (g_debug < 0) {
    const char* e = getenv("FLPR_GEMM_DEBUG");
    g_debug = e ? atoi(e) : 0;
  }
  return g_debug;
}

// lean epilogue: plain row-major bf16 / fp32 / fp32-atomic output, every 32-column chunk full and 16-byte aligned
static bool lean_ok(const GemmParams& p) {
  const long long esz = p.out_bf16 ? 2 : 4;
  const bool act = p.residual != nullptr || p.bias_n != nullptr || p.relu;
  return !p.trans_out && p.bias_m == nullptr && (!act || (p.out_bf16 && !p.atomic_add)) &&
         (reinterpret_cast<uintptr_t>(p.bias_n) % 16) == 0 && (reinterpret_cast<uintptr_t>(p.residual) % 16) == 0 &&
         (p.N % 32) == 0 && ((p.ldo * esz) % 16) == 0 && ((p.tap_stride * esz) % 16) == 0 &&
         (reinterpret_cast<uintptr_t>(p.out) % 16) == 0 && !(p.atomic_add && p.out_bf16) &&
         !(p.col_part != nullptr && !p.out_bf16) && !force_generic_epi();
}

// CTA-pair kernel eligibility: 256-wide tiles on the lean epilogue, enough pair tiles to fill the 74 SM pairs
static bool pair_ok(const GemmParams& p, int BN, int splits) {
  if (!use_pair() || BN != 256 || !lean_ok(p) || p.N < 256) return false;
  const long long t = (long long)((p.M + 255) / 256) * ((p.N + 255) / 256) * splits;
  return t >= 60;
}

template <int A_MODE, int B_MODE>
static int launch_pair(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p_in, int splits,
                       cudaStream_t st) {
  GemmParams p = p_in;
  p.debug = 0;
  p.tiles_m = (p.M + BM - 1) / BM;
  p.tiles_n = (p.N + PairLayout::BN - 1) / PairLayout::BN;
  p.tiles_z = splits;
  const int pair_tiles = ((p.tiles_m + 1) / 2) * p.tiles_n * splits;
  p.tiles_total = pair_tiles;
  auto kern = gemm_bf16_tcgen05_pair_kernel<A_MODE, B_MODE>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PairLayout::TOTAL);
    if (e != cudaSuccess) return set_err("cudaFuncSetAttribute(smem, pair)", (int)e);
    configured = true;
  }
  const int max_pairs = sm_count() / 2;
  const int pairs = pair_tiles < max_pairs ? pair_tiles : max_pairs;
  kern<<<2 * pairs, PairLayout::THREADS, PairLayout::TOTAL, st>>>(ta, tb, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_err(cudaGetErrorString(e), (int)e);
  return 0;
}

static int dispatch_pair(int a_mode, int b_mode, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                         int splits, cudaStream_t st) {
  if (a_mode == OP_KMAJOR && b_mode == OP_KMAJOR) return launch_pair<OP_KMAJOR, OP_KMAJOR>(ta, tb, p, splits, st);
  if (a_mode == OP_KMAJOR && b_mode == OP_MNMAJOR) return launch_pair<OP_KMAJOR, OP_MNMAJOR>(ta, tb, p, splits, st);
  if (a_mode == OP_MNMAJOR && b_mode == OP_KMAJOR) return launch_pair<OP_MNMAJOR, OP_KMAJOR>(ta, tb, p, splits, st);
  if (a_mode == OP_MNMAJOR && b_mode == OP_MNMAJOR) return launch_pair<OP_MNMAJOR, OP_MNMAJOR>(ta, tb, p, splits, st);
  if (a_mode == OP_CONV && b_mode == OP_KMAJOR) return launch_pair<OP_CONV, OP_KMAJOR>(ta, tb, p, splits, st);
  if (a_mode == OP_CONV && b_mode == OP_TAPFLIP) return launch_pair<OP_CONV, OP_TAPFLIP>(ta, tb, p, splits, st);
  if (a_mode == OP_MNMAJOR && b_mode == OP_CONV) return launch_pair<OP_MNMAJOR, OP_CONV>(ta, tb, p, splits, st);
  return set_err("unsupported operand mode combination (pair)", -3);
}

template <int BN, int A_MODE, int B_MODE>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p_in, int splits, cudaStream_t st) {
  GemmParams p = p_in;
  p.debug = debug_flags();
  p.tiles_m = (p.M + BM - 1) / BM;
  p.tiles_n = (p.N + BN - 1) / BN;
  p.tiles_z = splits;
  p.tiles_total = p.tiles_m * p.tiles_n * splits;
  if (use_persist()) {
    using L = PersistLayout<BN>;
    const bool lean = lean_ok(p);
    const int slots = sm_count() * L::MIN_CTAS;
    const int grid = p.tiles_total < slots ? p.tiles_total : slots;
    if (lean) {
      auto kern = gemm_bf16_tcgen05_persistent_kernel<BN, A_MODE, B_MODE, EPI_LEAN>;
      static bool configured = false;
      if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
        if (e != cudaSuccess) return set_err("cudaFuncSetAttribute(smem, persistent lean)", (int)e);
        configured = true;
      }
      kern<<<grid, L::THREADS, L::TOTAL, st>>>(ta, tb, p);
    } else {
      auto kern = gemm_bf16_tcgen05_persistent_kernel<BN, A_MODE, B_MODE, EPI_GENERIC>;
      static bool configured = false;
      if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
        if (e != cudaSuccess) return set_err("cudaFuncSetAttribute(smem, persistent)", (int)e);
        configured = true;
      }
      kern<<<grid, L::THREADS, L::TOTAL, st>>>(ta, tb, p);
    }
  } else {
    using L = SmemLayout<BN>;
    auto kern = gemm_bf16_tcgen05_kernel<BN, A_MODE, B_MODE>;
    static bool configured = false;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
      if (e != cudaSuccess) return set_err("cudaFuncSetAttribute(smem)", (int)e);
      configured = true;
    }
    dim3 grid(p.tiles_n, p.tiles_m, splits);
    kern<<<grid, 256, L::TOTAL, st>>>(ta, tb, p);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_err(cudaGetErrorString(e), (int)e);
  return 0;
}

template <int BN>
static int dispatch_modes(int a_mode, int b_mode, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                          int splits, cudaStream_t st) {
  if (a_mode == OP_KMAJOR && b_mode == OP_KMAJOR) return launch<BN, OP_KMAJOR, OP_KMAJOR>(ta, tb, p, splits, st);
  if (a_mode == OP_KMAJOR && b_mode == OP_MNMAJOR) return launch<BN, OP_KMAJOR, OP_MNMAJOR>(ta, tb, p, splits, st);
  if (a_mode == OP_MNMAJOR && b_mode == OP_KMAJOR) return launch<BN, OP_MNMAJOR, OP_KMAJOR>(ta, tb, p, splits, st);
  if (a_mode == OP_MNMAJOR && b_mode == OP_MNMAJOR) return launch<BN, OP_MNMAJOR, OP_MNMAJOR>(ta, tb, p, splits, st);
  if (a_mode == OP_CONV && b_mode == OP_KMAJOR) return launch<BN, OP_CONV, OP_KMAJOR>(ta, tb, p, splits, st);
  if (a_mode == OP_CONV && b_mode == OP_TAPFLIP) return launch<BN, OP_CONV, OP_TAPFLIP>(ta, tb, p, splits, st);
  if (a_mode == OP_MNMAJOR && b_mode == OP_CONV) return launch<BN, OP_MNMAJOR, OP_CONV>(ta, tb, p, splits, st);
  return set_err("unsupported operand mode combination", -3);
}

static int dispatch_bn(int BN, int a_mode, int b_mode, const CUtensorMap& ta, const CUtensorMap& tb,
                       const GemmParams& p, int splits, cudaStream_t st, bool pair = false) {
  if (pair) return dispatch_pair(a_mode, b_mode, ta, tb, p, splits, st);
  if (BN == 64) return dispatch_modes<64>(a_mode, b_mode, ta, tb, p, splits, st);
  if (BN == 128) return dispatch_modes<128>(a_mode, b_mode, ta, tb, p, splits, st);
  return dispatch_modes<256>(a_mode, b_mode, ta, tb, p, splits, st);
}

// Tile width. 128 x 256 tiles halve the B-operand smem traffic per MMA (a 128 x 128 tile is smem-bandwidth bound at
// ~50 % of the tensor peak) but leave one CTA per SM, so they are used when they still give ~a full wave.
static int pick_bn(int M, int N, int z, int requested) {
  if (requested == 64 || requested == 128 || requested == 256) return requested;
  if (N <= 64) return 64;
  if (!use_persist()) return 128;
  const long long mt = (M + BM - 1) / BM;
  const long long t256 = mt * ((N + 255) / 256) * (z > 0 ? z : 1);
  if (N >= 256 && t256 >= 120) return 256;
  return 128;
}

}  // namespace flpr

using namespace flpr;

extern "C" {

const char* flpr_gemm_last_error() { return g_err; }

// 1: persistent kernel (default), 0: classic one-tile-per-CTA kernel, -1: re-read FLPR_GEMM_PERSIST.
void flpr_gemm_set_persistent(int on) { g_persist = on; }

// 1: use the cta_group::2 (CTA-pair) kernel where eligible, 0: never, -1: re-read FLPR_GEMM_2CTA.
void flpr_gemm_set_pair(int on) { g_pair = on; }

// bottleneck-isolation switches of scripts/gemm_*: debug bit mask (see GemmParams::debug) and forced generic epilogue;
// -1 re-reads the FLPR_GEMM_DEBUG / FLPR_GEMM_GENERIC_EPI environment variables at the next launch.
void flpr_gemm_set_debug(int flags) { g_debug = flags; }
void flpr_gemm_set_generic_epilogue(int on) { g_generic_epi = on; }

// D = alpha * op(A) * op(B)^T.  a_mode/b_mode: 0 = [rows,K] (ld = row stride), 1 = [K,rows] (ld = K-row stride).
// col_part (optional, fp32 [ceil(M/128)*4][2][N]): per-32-row partial column sums / sums of squares of the fp32
// result (fused batch-norm statistics); requires split_k <= 1 and no transposed output.
int flpr_gemm_bf16(const void* A, const void* B, void* out, int M, int N, int K, long long lda, long long ldb,
                   long long ldo, int a_mode, int b_mode, int out_bf16, int trans_out, float alpha,
                   const float* bias_n, const float* bias_m, int relu, const void* residual, int split_k, int bn_req,
                   float* col_part, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  bind_device_of(A);
  if ((lda % 8) || (ldb % 8)) return set_err("lda/ldb must be multiples of 8 elements (16B TMA stride)", -4);
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15))
    return set_err("A/B must be 16B aligned", -5);
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K;
  p.kb_total = (K + BK - 1) / BK;
  int splits = split_k > 1 ? split_k : 1;
  if (splits > p.kb_total) splits = p.kb_total;
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  const int BN = pick_bn(M, N, splits, bn_req);
  p.out = out; p.ldo = ldo; p.out_bf16 = out_bf16; p.trans_out = trans_out;
  p.atomic_add = splits > 1 ? 1 : 0;
  if (p.atomic_add && out_bf16) return set_err("split-K requires fp32 output", -6);
  if (col_part != nullptr && (p.atomic_add || trans_out)) return set_err("col_part needs split_k=1, no transpose", -12);
  p.col_part = col_part;
  p.alpha = alpha; p.bias_n = bias_n; p.bias_m = bias_m; p.relu = relu;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  const bool pair = pair_ok(p, BN, splits);
  CUtensorMap ta, tb;
  int rc;
  {
    uint64_t dims[2], str[1];
    uint32_t box[2];
    if (a_mode == OP_KMAJOR) {
      dims[0] = (uint64_t)K; dims[1] = (uint64_t)M; box[0] = BK; box[1] = BM;
    } else {
      dims[0] = (uint64_t)M; dims[1] = (uint64_t)K; box[0] = 64; box[1] = BK;
    }
    str[0] = (uint64_t)lda * 2;
    if ((rc = get_map(&ta, A, 2, dims, str, box))) return rc;
    if (b_mode == OP_KMAJOR) {
      dims[0] = (uint64_t)K; dims[1] = (uint64_t)N; box[0] = BK; box[1] = (uint32_t)(pair ? PairLayout::BNH : BN);
    } else {
      dims[0] = (uint64_t)N; dims[1] = (uint64_t)K; box[0] = 64; box[1] = BK;
    }
    str[0] = (uint64_t)ldb * 2;
    if ((rc = get_map(&tb, B, 2, dims, str, box))) return rc;
  }
  return dispatch_bn(BN, a_mode, b_mode, ta, tb, p, splits, stream, pair);
}

static int conv_tiling(int H, int W, int* TH, int* NB, int* tiles_per_img) {
  if (W > 128 || (128 % W)) return set_err("conv: W must divide 128", -8);
  if (H * W <= 128) {
    if (128 % (H * W)) return set_err("conv: H*W must divide 128", -9);
    *TH = H; *NB = 128 / (H * W); *tiles_per_img = 1;
  } else {
    *TH = 128 / W;
    if (H % *TH) return set_err("conv: 128/W must divide H", -10);
    *NB = 1; *tiles_per_img = H / *TH;
  }
  return 0;
}

// Implicit-GEMM convolution (stride 1 or 2): X [NIMG,H,W,C] bf16 NHWC, Wt [Cout, KH*KW*C] bf16 (tap-major, then C),
// out [NIMG*Ho*Wo, Cout]. Requires C % 64 == 0, Wo a power of two <= 128 and (Ho*Wo) | 128 or 128/Wo | Ho.
// A strided convolution is the same pipeline: the TMA tensor map walks the input with element strides {1,s,s,1}.
// col_part: see flpr_gemm_bf16 (fused batch-norm statistics of the conv output).
int flpr_conv_nhwc_bf16(const void* X, const void* Wt, void* out, int NIMG, int H, int W, int C, int Cout, int KH,
                        int KW, int pad_h, int pad_w, int out_bf16, float alpha, const float* bias_n, int relu,
                        const void* residual, int bn_req, float* col_part, int stride, long long x_stride_w,
                        long long x_stride_h, long long x_stride_n, cudaStream_t stream) {
  // x_stride_{w,h,n}: element strides of X (0 = dense NHWC). A W stride smaller than C makes the per-pixel "channel"
  // vectors overlap: that is how the 7x7 stem convolution runs on 4-cell windows of a space-to-depth input.
  bind_device_of(X);
  if (C % 64) return set_err("conv: C must be a multiple of 64", -7);
  if (stride != 1 && stride != 2) return set_err("conv: stride must be 1 or 2", -13);
  const int Ho = (H + 2 * pad_h - KH) / stride + 1;
  const int Wo = (W + 2 * pad_w - KW) / stride + 1;
  int TH, NB, tiles_per_img, rc;
  if ((rc = conv_tiling(Ho, Wo, &TH, &NB, &tiles_per_img))) return rc;
  const int M = NIMG * Ho * Wo;
  const int K = KH * KW * C;
  const int BN = pick_bn(M, Cout, 1, bn_req);
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = Cout; p.K = K;
  p.kb_total = K / BK; p.kb_per_split = p.kb_total;
  p.out = out; p.ldo = Cout; p.out_bf16 = out_bf16; p.alpha = alpha; p.bias_n = bias_n; p.relu = relu;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.col_part = col_part;
  p.cH = Ho; p.cW = Wo; p.cC = C; p.cTH = TH; p.cNB = NB; p.cKW = KW; p.cPadH = pad_h; p.cPadW = pad_w;
  p.cTilesPerImg = tiles_per_img; p.cTaps = KH * KW; p.cStride = stride;
  const bool pair = pair_ok(p, BN, 1);
  CUtensorMap ta, tb;
  {
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)NIMG};
    uint64_t str[3] = {(uint64_t)(x_stride_w ? x_stride_w : C) * 2,
                       (uint64_t)(x_stride_h ? x_stride_h : (long long)W * C) * 2,
                       (uint64_t)(x_stride_n ? x_stride_n : (long long)H * W * C) * 2};
    uint32_t box[4] = {64, (uint32_t)(Wo * stride), (uint32_t)(TH * stride), (uint32_t)NB};
    if ((rc = get_map(&ta, X, 4, dims, str, box, (uint32_t)stride))) return rc;
    uint64_t d2[2] = {(uint64_t)K, (uint64_t)Cout};
    uint64_t s2[1] = {(uint64_t)K * 2};
    uint32_t b2[2] = {BK, (uint32_t)(pair ? PairLayout::BNH : BN)};
    if ((rc = get_map(&tb, Wt, 2, d2, s2, b2))) return rc;
  }
  return dispatch_bn(BN, OP_CONV, OP_KMAJOR, ta, tb, p, 1, stream, pair);
}

// Data gradient of a stride-1 convolution, straight from the FORWARD weight Wt [Cout, KH*KW*Cin] (no flipped /
// transposed copy): dX[NIMG*H*W, Cin] = sum_taps dY(shifted by the mirrored tap)[., Cout] x Wt[:, tap, :].
// dY: [NIMG,H,W,Cout] bf16. pad_h/pad_w are the FORWARD paddings (dgrad pads with K-1-pad).
int flpr_conv_dgrad_nhwc_bf16(const void* dY, const void* Wt, void* out, int NIMG, int H, int W, int Cin, int Cout,
                              int KH, int KW, int pad_h, int pad_w, int out_bf16, int bn_req, cudaStream_t stream) {
  bind_device_of(dY);
  if (Cout % 64) return set_err("conv dgrad: Cout must be a multiple of 64", -7);
  if (Cin % 64) return set_err("conv dgrad: Cin must be a multiple of 64", -7);
  int TH, NB, tiles_per_img, rc;
  if ((rc = conv_tiling(H, W, &TH, &NB, &tiles_per_img))) return rc;
  const int M = NIMG * H * W;
  const int K = KH * KW * Cout;
  const int BN = pick_bn(M, Cin, 1, bn_req);
  CUtensorMap ta, tb;
  {
    uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)NIMG};
    uint64_t str[3] = {(uint64_t)Cout * 2, (uint64_t)W * Cout * 2, (uint64_t)H * W * Cout * 2};
    uint32_t box[4] = {64, (uint32_t)W, (uint32_t)TH, (uint32_t)NB};
    if ((rc = get_map(&ta, dY, 4, dims, str, box))) return rc;
    uint64_t d2[2] = {(uint64_t)KH * KW * Cin, (uint64_t)Cout};
    uint64_t s2[1] = {(uint64_t)KH * KW * Cin * 2};
    uint32_t b2[2] = {64, BK};
    if ((rc = get_map(&tb, Wt, 2, d2, s2, b2))) return rc;
  }
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = Cin; p.K = K;
  p.kb_total = K / BK; p.kb_per_split = p.kb_total;
  p.out = out; p.ldo = Cin; p.out_bf16 = out_bf16; p.alpha = 1.f;
  p.cH = H; p.cW = W; p.cC = Cout; p.cTH = TH; p.cNB = NB; p.cKW = KW;
  p.cPadH = KH - 1 - pad_h; p.cPadW = KW - 1 - pad_w;
  p.cTilesPerImg = tiles_per_img; p.cTaps = KH * KW; p.cStride = 1;
  return dispatch_bn(BN, OP_CONV, OP_TAPFLIP, ta, tb, p, 1, stream, pair_ok(p, BN, 1));
}

// Weight gradient of a stride-1 convolution: out[Cout, KH*KW*C] (fp32, tap-major then C) +=
//   sum over pixels dY[p, cout] * Xshift_tap[p, c].  One launch covers every tap (grid.z = taps * split_k).
// dY: [NIMG*H*W, Cout] bf16, X: [NIMG,H,W,C] bf16. `out` must be zeroed when split_k > 1 (atomic accumulation).
int flpr_conv_wgrad_nhwc_bf16(const void* X, const void* dY, float* out, int NIMG, int H, int W, int C, int Cout,
                              int KH, int KW, int pad_h, int pad_w, int split_k, int bn_req, cudaStream_t stream) {
  bind_device_of(X);
  if (C % 64) return set_err("conv wgrad: C must be a multiple of 64", -7);
  if (Cout % 8) return set_err("conv wgrad: Cout must be a multiple of 8", -11);
  if (W > 64 || (64 % W)) return set_err("conv wgrad: W must divide 64", -8);
  int TH, NB;
  if (H * W >= 64) {
    if ((H * W) % 64) return set_err("conv wgrad: H*W must be a multiple of 64", -9);
    TH = 64 / W; NB = 1;
  } else {
    if (64 % (H * W)) return set_err("conv wgrad: H*W must divide 64", -9);
    TH = H; NB = 64 / (H * W);
  }
  const int M = Cout, N = C, K = NIMG * H * W;
  const int BN = (bn_req == 64 || bn_req == 128 || bn_req == 256) ? bn_req : (N >= 128 ? 128 : 64);
  CUtensorMap ta, tb;
  int rc;
  {
    uint64_t d2[2] = {(uint64_t)M, (uint64_t)K};
    uint64_t s2[1] = {(uint64_t)Cout * 2};
    uint32_t b2[2] = {64, BK};
    if ((rc = get_map(&ta, dY, 2, d2, s2, b2))) return rc;
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)NIMG};
    uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    uint32_t box[4] = {64, (uint32_t)W, (uint32_t)TH, (uint32_t)NB};
    if ((rc = get_map(&tb, X, 4, dims, str, box))) return rc;
  }
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K;
  p.kb_total = (K + BK - 1) / BK;
  int splits = split_k > 1 ? split_k : 1;
  if (splits > p.kb_total) splits = p.kb_total;
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  p.out = out; p.ldo = (long long)KH * KW * C; p.out_bf16 = 0; p.alpha = 1.f;
  p.atomic_add = splits > 1 ? 1 : 0;
  p.cH = H; p.cW = W; p.cC = C; p.cTH = TH; p.cNB = NB; p.cKW = KW; p.cPadH = pad_h; p.cPadW = pad_w;
  p.cSplits = splits; p.tap_stride = C; p.cTaps = KH * KW;
  const int gz = KH * KW * splits;
  return dispatch_bn(BN, OP_MNMAJOR, OP_CONV, ta, tb, p, gz, stream, pair_ok(p, BN, gz));
}

}  // extern "C"
