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
#include <string.h>

#include <mutex>
#include <unordered_map>

#include "ptx.cuh"

namespace flpr {

enum { OP_KMAJOR = 0, OP_MNMAJOR = 1, OP_CONV = 2 };

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
        const int kb = kb_begin + i;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * L::STAGE_BYTES;
        uint8_t* sb = sa + A_STAGE_BYTES;
        mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
        if constexpr (A_MODE == OP_KMAJOR) {
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m0);
        } else if constexpr (A_MODE == OP_MNMAJOR) {
          tma_load_2d(sa, &tmA, &full_bar[stage], m0, kb * BK);
          tma_load_2d(sa + 64 * 128, &tmA, &full_bar[stage], m0 + 64, kb * BK);
        } else {
          const int chunks = p.cC / BK;
          const int tap = kb / chunks;
          const int cc = kb - tap * chunks;
          const int kh = tap / p.cKW;
          const int kw = tap - kh * p.cKW;
          const int mt = blockIdx.y;
          int img0, h0;
          if (p.cTilesPerImg <= 1) {
            img0 = mt * p.cNB;
            h0 = 0;
          } else {
            img0 = mt / p.cTilesPerImg;
            h0 = (mt - img0 * p.cTilesPerImg) * p.cTH;
          }
          tma_load_4d(sa, &tmA, &full_bar[stage], cc * BK, kw - p.cPadW, h0 + kh - p.cPadH, img0);
        }
        if constexpr (B_MODE == OP_KMAJOR) {
          tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, n0);
        } else if constexpr (B_MODE == OP_MNMAJOR) {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * 64 * 128, &tmB, &full_bar[stage], n0 + j * 64, kb * BK);
        } else {
          // weight-gradient of a convolution: K runs over pixels (64 per block), the filter tap (blockIdx.z / splits)
          // shifts the 4-D box, zero padding comes from TMA out-of-bounds fill
          const int tap = blockIdx.z / p.cSplits;
          const int kh = tap / p.cKW;
          const int kw = tap - kh * p.cKW;
          const int p0 = kb * BK;
          const int hw = p.cH * p.cW;
          const int img = p0 / hw;
          const int h0 = (p0 - img * hw) / p.cW;
#pragma unroll
          for (int j = 0; j < BN / 64; ++j)
            tma_load_4d(sb + j * 64 * 128, &tmB, &full_bar[stage], n0 + j * 64, kw - p.cPadW, h0 + kh - p.cPadH, img);
        }
        if (++stage == L::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp_id == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MODE == OP_MNMAJOR, B_MODE != OP_KMAJOR);
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < num_kb; ++i) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
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
          umma_f16(tmem_base, da, db, idesc, (i | k) != 0);
        }
      }
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
    const int row = m0 + q * 32 + lane;
    if (num_kb > 0) {
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
    }
    const float bm = (p.bias_m != nullptr && row < p.M) ? p.bias_m[row] : 0.f;
    const long long tap_off = (B_MODE == OP_CONV) ? (long long)(blockIdx.z / p.cSplits) * p.tap_stride : 0;
    // bf16 row-major output without residual: stores are staged through smem (pipeline stage 0 is free once the
    // accumulator barrier has fired) so that each store instruction writes 8 full 64-byte row segments
    const bool use_stage = p.out_bf16 && !p.trans_out && !p.atomic_add && p.residual == nullptr &&
                           ((p.ldo & 7) == 0) && (((n0 + tap_off) & 7) == 0) &&
                           ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
    uint8_t* stage_buf = smem + q * (32 * 80);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      if (num_kb > 0) {
        tmem_ld_32x32b_x32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(c0), r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
      if (n0 + c0 >= p.N) continue;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float x = __uint_as_float(r[j]) * p.alpha + bm;
        const int col = n0 + c0 + j;
        if (p.bias_n != nullptr && col < p.N) x += p.bias_n[col];
        v[j] = x;
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
        if (p.residual != nullptr) {
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
              uint8_t* st = stage_buf + lane * 80;
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) *reinterpret_cast<uint4*>(st + j4 * 16) = pk[j4];
              __syncwarp();
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int rr = 8 * j + (lane >> 2);
                const uint4 u = *reinterpret_cast<const uint4*>(stage_buf + rr * 80 + (lane & 3) * 16);
                const int grow = m0 + q * 32 + rr;
                if (grow < p.M)
                  *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + (long long)grow * p.ldo + n0 +
                                            c0 + tap_off + (lane & 3) * 8) = u;
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
  }

  tc_fence_before();
  __syncthreads();
  if (warp_id == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
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
  uint32_t b0, b1, b2, b3, rank;
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
                   const uint32_t* box) {
  MapKey k;
  memset(&k, 0, sizeof(k));
  k.ptr = ptr;
  k.rank = rank;
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

template <int BN, int A_MODE, int B_MODE>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int splits, cudaStream_t st) {
  using L = SmemLayout<BN>;
  auto kern = gemm_bf16_tcgen05_kernel<BN, A_MODE, B_MODE>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    if (e != cudaSuccess) return set_err("cudaFuncSetAttribute(smem)", (int)e);
    configured = true;
  }
  dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, splits);
  kern<<<grid, 256, L::TOTAL, st>>>(ta, tb, p);
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
  if (a_mode == OP_MNMAJOR && b_mode == OP_CONV) return launch<BN, OP_MNMAJOR, OP_CONV>(ta, tb, p, splits, st);
  return set_err("unsupported operand mode combination", -3);
}

static int pick_bn(int M, int N, int requested) {
  if (requested == 64 || requested == 128 || requested == 256) return requested;
  // Aim for >= ~1 wave of 148 SMs; prefer the widest tile that still fills the machine.
  const long long mt = (M + BM - 1) / BM;
  (void)mt;
  if (N <= 64) return 64;
  return 128;   // two resident CTAs per SM hide the epilogue; 256-wide tiles only on request
}

}  // namespace flpr

using namespace flpr;

extern "C" {

const char* flpr_gemm_last_error() { return g_err; }

// D = alpha * op(A) * op(B)^T.  a_mode/b_mode: 0 = [rows,K] (ld = row stride), 1 = [K,rows] (ld = K-row stride).
int flpr_gemm_bf16(const void* A, const void* B, void* out, int M, int N, int K, long long lda, long long ldb,
                   long long ldo, int a_mode, int b_mode, int out_bf16, int trans_out, float alpha,
                   const float* bias_n, const float* bias_m, int relu, const void* residual, int split_k, int bn_req,
                   cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  bind_device_of(A);
  if ((lda % 8) || (ldb % 8)) return set_err("lda/ldb must be multiples of 8 elements (16B TMA stride)", -4);
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15))
    return set_err("A/B must be 16B aligned", -5);
  const int BN = pick_bn(M, N, bn_req);
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
      dims[0] = (uint64_t)K; dims[1] = (uint64_t)N; box[0] = BK; box[1] = (uint32_t)BN;
    } else {
      dims[0] = (uint64_t)N; dims[1] = (uint64_t)K; box[0] = 64; box[1] = BK;
    }
    str[0] = (uint64_t)ldb * 2;
    if ((rc = get_map(&tb, B, 2, dims, str, box))) return rc;
  }
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K;
  p.kb_total = (K + BK - 1) / BK;
  int splits = split_k > 1 ? split_k : 1;
  if (splits > p.kb_total) splits = p.kb_total;
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  p.out = out; p.ldo = ldo; p.out_bf16 = out_bf16; p.trans_out = trans_out;
  p.atomic_add = splits > 1 ? 1 : 0;
  if (p.atomic_add && out_bf16) return set_err("split-K requires fp32 output", -6);
  p.alpha = alpha; p.bias_n = bias_n; p.bias_m = bias_m; p.relu = relu;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  if (BN == 64) return dispatch_modes<64>(a_mode, b_mode, ta, tb, p, splits, stream);
  if (BN == 128) return dispatch_modes<128>(a_mode, b_mode, ta, tb, p, splits, stream);
  return dispatch_modes<256>(a_mode, b_mode, ta, tb, p, splits, stream);
}

// Implicit-GEMM convolution, stride 1: X [NIMG,H,W,C] bf16 NHWC, Wt [Cout, KH*KW*C] bf16 (tap-major, then C),
// out [NIMG*H*W, Cout]. Requires C % 64 == 0, W a power of two <= 128 and (H*W) | 128 or 128/W | H.
int flpr_conv_nhwc_bf16(const void* X, const void* Wt, void* out, int NIMG, int H, int W, int C, int Cout, int KH,
                        int KW, int pad_h, int pad_w, int out_bf16, float alpha, const float* bias_n, int relu,
                        const void* residual, int bn_req, cudaStream_t stream) {
  bind_device_of(X);
  if (C % 64) return set_err("conv: C must be a multiple of 64", -7);
  if (W > 128 || (128 % W)) return set_err("conv: W must divide 128", -8);
  int TH, NB, tiles_per_img;
  if (H * W <= 128) {
    if (128 % (H * W)) return set_err("conv: H*W must divide 128", -9);
    TH = H; NB = 128 / (H * W); tiles_per_img = 1;
  } else {
    TH = 128 / W;
    if (H % TH) return set_err("conv: 128/W must divide H", -10);
    NB = 1; tiles_per_img = H / TH;
  }
  const int M = NIMG * H * W;
  const int K = KH * KW * C;
  const int BN = pick_bn(M, Cout, bn_req);
  CUtensorMap ta, tb;
  int rc;
  {
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)NIMG};
    uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    uint32_t box[4] = {64, (uint32_t)W, (uint32_t)TH, (uint32_t)NB};
    if ((rc = get_map(&ta, X, 4, dims, str, box))) return rc;
    uint64_t d2[2] = {(uint64_t)K, (uint64_t)Cout};
    uint64_t s2[1] = {(uint64_t)K * 2};
    uint32_t b2[2] = {BK, (uint32_t)BN};
    if ((rc = get_map(&tb, Wt, 2, d2, s2, b2))) return rc;
  }
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = Cout; p.K = K;
  p.kb_total = K / BK; p.kb_per_split = p.kb_total;
  p.out = out; p.ldo = Cout; p.out_bf16 = out_bf16; p.alpha = alpha; p.bias_n = bias_n; p.relu = relu;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.cH = H; p.cW = W; p.cC = C; p.cTH = TH; p.cNB = NB; p.cKW = KW; p.cPadH = pad_h; p.cPadW = pad_w;
  p.cTilesPerImg = tiles_per_img;
  if (BN == 64) return dispatch_modes<64>(OP_CONV, OP_KMAJOR, ta, tb, p, 1, stream);
  if (BN == 128) return dispatch_modes<128>(OP_CONV, OP_KMAJOR, ta, tb, p, 1, stream);
  return dispatch_modes<256>(OP_CONV, OP_KMAJOR, ta, tb, p, 1, stream);
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
  p.cSplits = splits; p.tap_stride = C;
  const int gz = KH * KW * splits;
  if (BN == 64) return dispatch_modes<64>(OP_MNMAJOR, OP_CONV, ta, tb, p, gz, stream);
  if (BN == 128) return dispatch_modes<128>(OP_MNMAJOR, OP_CONV, ta, tb, p, gz, stream);
  return dispatch_modes<256>(OP_MNMAJOR, OP_CONV, ta, tb, p, gz, stream);
}

}  // extern "C"
