// Host-side SIMT emulator for the flpr CUDA-core kernels (tests only; no GPU, no CUDA runtime call).
//
// A kernel source is compiled as plain C++ (tests/emu/build_emu.py rewrites `k<<<grid, block, smem, stream>>>(args)` into
// flpr_emu::launch(...) and includes this header instead of csrc/ptx.cuh). Every CUDA thread of a block is a ucontext
// fiber on ONE OS thread; blocks run one after the other:
//   * __syncthreads / __syncthreads_or : cooperative block barrier (threads that returned from the kernel stop counting,
//                                        as on the hardware);
//   * __shfl_xor_sync / __shfl_sync / __shfl_down_sync : exchange through a per-warp buffer between two warp barriers -
//                                        the lanes really read each other's registers, so a wrong lane mask or a shuffle
//                                        under divergent control flow shows up as a wrong result or a reported deadlock;
//   * __shared__                       : a function-local static (one block at a time).
// What this checks: index mathematics, reductions, argument marshalling of the extern "C" entry points, barrier placement
// (a barrier that not every live thread reaches is reported as a deadlock instead of hanging). What it cannot check:
// alignment faults, memory-model races between warps, performance.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <functional>
#include <vector>

#undef __shared__
#define __shared__ static
#undef __global__
#define __global__
#undef __device__
#define __device__
#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __forceinline__
#define __forceinline__ inline

namespace flpr_emu {

struct Dim3 {
  unsigned x = 0, y = 0, z = 0;
};
static Dim3 g_tid, g_bid, g_bdim, g_gdim;   // (internal linkage: one emulator state per emulated library)

constexpr int MAX_THREADS = 1024;
constexpr int MAX_WARPS = MAX_THREADS / 32;
constexpr size_t STACK_BYTES = 64 * 1024;

struct Block {
  int n = 0, alive = 0, cur = 0;
  ucontext_t main_ctx;
  ucontext_t ctx[MAX_THREADS];
  bool done[MAX_THREADS];
  // block barrier
  int sync_count = 0, sync_or = 0, sync_or_result = 0;
  unsigned sync_gen = 0;
  // warp barriers + exchange buffers
  int warp_alive[MAX_WARPS], warp_count[MAX_WARPS];
  unsigned warp_gen[MAX_WARPS];
  uint64_t warp_buf[MAX_WARPS][32];
  const std::function<void()>* body = nullptr;
  unsigned long long progress = 0;   // bumped whenever a barrier releases or a thread exits (deadlock detection)
};
static Block g_blk;
static char* g_stacks = nullptr;
static int g_deadlocks = 0;

inline void yield() { swapcontext(&g_blk.ctx[g_blk.cur], &g_blk.main_ctx); }

inline void release_block_barrier() {
  Block& b = g_blk;
  b.sync_or_result = b.sync_or;
  b.sync_or = 0;
  b.sync_count = 0;
  b.sync_gen++;
  b.progress++;
}

inline int sync_impl(int pred) {
  Block& b = g_blk;
  const unsigned g = b.sync_gen;
  b.sync_or |= (pred != 0);
  if (++b.sync_count == b.alive)
    release_block_barrier();
  else
    while (b.sync_gen == g) yield();
  return b.sync_or_result;
}

inline void warp_barrier(int w) {
  Block& b = g_blk;
  const unsigned g = b.warp_gen[w];
  if (++b.warp_count[w] == b.warp_alive[w]) {
    b.warp_count[w] = 0;
    b.warp_gen[w]++;
    b.progress++;
  } else {
    while (b.warp_gen[w] == g) yield();
  }
}

inline uint64_t exchange(uint64_t v, int src_lane) {
  Block& b = g_blk;
  const int w = b.cur >> 5, l = b.cur & 31;
  b.warp_buf[w][l] = v;
  warp_barrier(w);
  const uint64_t r = b.warp_buf[w][src_lane & 31];
  warp_barrier(w);
  return r;
}

template <class T>
inline T shfl_from(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  uint64_t bits = 0;
  memcpy(&bits, &v, sizeof(T));
  bits = exchange(bits, src_lane);
  T out;
  memcpy(&out, &bits, sizeof(T));
  return out;
}

inline void fiber_entry() {
  Block& b = g_blk;
  (*b.body)();
  const int t = b.cur, w = t >> 5;
  b.done[t] = true;
  b.alive--;
  b.warp_alive[w]--;
  b.progress++;
  // threads that exited no longer take part in barriers: release the ones that were only waiting for this thread
  if (b.sync_count > 0 && b.sync_count == b.alive) release_block_barrier();
  if (b.warp_count[w] > 0 && b.warp_count[w] == b.warp_alive[w]) {
    b.warp_count[w] = 0;
    b.warp_gen[w]++;
  }
  // uc_link returns to main_ctx
}

inline void run_block(unsigned bid, unsigned grid, unsigned block, const std::function<void()>& body) {
  Block& b = g_blk;
  if (block == 0 || block > (unsigned)MAX_THREADS) {
    fprintf(stderr, "[flpr_emu] unsupported block size %u\n", block);
    abort();
  }
  if (g_stacks == nullptr) g_stacks = (char*)malloc(STACK_BYTES * MAX_THREADS);
  b.n = b.alive = (int)block;
  b.body = &body;
  b.sync_count = b.sync_or = b.sync_or_result = 0;
  b.sync_gen = 0;
  for (int w = 0; w < MAX_WARPS; ++w) {
    b.warp_count[w] = 0;
    b.warp_gen[w] = 0;
    const int lo = w * 32;
    b.warp_alive[w] = (int)block > lo ? ((int)block - lo >= 32 ? 32 : (int)block - lo) : 0;
  }
  g_bid.x = bid;
  g_gdim.x = grid;
  g_bdim.x = block;
  g_bid.y = g_bid.z = 0;
  g_gdim.y = g_gdim.z = g_bdim.y = g_bdim.z = 1;
  for (unsigned t = 0; t < block; ++t) {
    b.done[t] = false;
    getcontext(&b.ctx[t]);
    b.ctx[t].uc_stack.ss_sp = g_stacks + STACK_BYTES * t;
    b.ctx[t].uc_stack.ss_size = STACK_BYTES;
    b.ctx[t].uc_link = &b.main_ctx;
    makecontext(&b.ctx[t], (void (*)())fiber_entry, 0);
  }
  unsigned long long last = ~0ull;
  int idle_rounds = 0;
  while (b.alive > 0) {
    for (unsigned t = 0; t < block; ++t) {
      if (b.done[t]) continue;
      b.cur = (int)t;
      g_tid.x = t;
      g_tid.y = g_tid.z = 0;
      swapcontext(&b.main_ctx, &b.ctx[t]);
    }
    if (b.progress == last) {
      if (++idle_rounds > 4) {   // every live thread is parked at a barrier that can never release
        fprintf(stderr, "[flpr_emu] deadlock in block %u: %d threads wait at a barrier not every live thread reaches\n",
                bid, b.alive);
        g_deadlocks++;
        return;
      }
    } else {
      idle_rounds = 0;
      last = b.progress;
    }
  }
}

inline void launch(unsigned grid, unsigned block, const std::function<void()>& body) {
  for (unsigned bid = 0; bid < grid; ++bid) run_block(bid, grid, block, body);
}

}  // namespace flpr_emu

#define threadIdx flpr_emu::g_tid
#define blockIdx flpr_emu::g_bid
#define blockDim flpr_emu::g_bdim
#define gridDim flpr_emu::g_gdim

inline void __syncthreads() { flpr_emu::sync_impl(0); }
inline int __syncthreads_or(int pred) { return flpr_emu::sync_impl(pred); }
inline void __syncwarp(unsigned = 0xffffffffu) { flpr_emu::warp_barrier(flpr_emu::g_blk.cur >> 5); }

template <class T>
inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
  return flpr_emu::shfl_from(v, (flpr_emu::g_blk.cur & 31) ^ lane_mask);
}
template <class T>
inline T __shfl_sync(unsigned, T v, int src_lane) {
  return flpr_emu::shfl_from(v, src_lane);
}
template <class T>
inline T __shfl_down_sync(unsigned, T v, unsigned delta) {
  const int l = flpr_emu::g_blk.cur & 31;
  const T r = flpr_emu::shfl_from(v, l + (int)delta);
  return (l + (int)delta < 32) ? r : v;
}

inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }

// ---- the pieces of csrc/ptx.cuh the CUDA-core kernels use -----------------------------------------------------------------
namespace flpr {
inline void bind_device_of(const void*) {}
inline float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
inline float warp_max(float v) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
}  // namespace flpr
using flpr::bind_device_of;

#define cudaGetLastError() (cudaSuccess)

extern "C" int flpr_emu_deadlocks() { return flpr_emu::g_deadlocks; }
