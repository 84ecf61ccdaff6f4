"""The CUDA-core kernels of ``csrc/layer_ops.cu`` executed on the CPU under a SIMT emulator (``tests/emu``).

Those kernels (weight composition, Swin token kernels, LayerNorm forward / backward, dispatch apply) were written after
the builder's last GPU session. Here the *kernel source itself* is compiled for the host - every CUDA thread a fiber,
``__syncthreads`` / warp shuffles / ``__shared__`` emulated - and driven through the same ctypes wrappers and ``extern
"C"`` entry points as on the device, with CPU tensors. Each check compares against a plain fp32 PyTorch formulation
(the bodies are shared with ``test_cpu_late.py`` / ``test_zz_gpu_late.py``). Covered: index mathematics, reductions,
barrier placement (a barrier not every live thread reaches is reported as a deadlock), argument marshalling. Not
covered: alignment faults, cross-warp memory-model races, performance - the on-device self-check gate remains."""
import shutil

import pytest
import torch

import late_checks as L


@pytest.fixture(scope="module")
def emu():
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    from emu.build_emu import build
    from flpr_b200.ops import layer as lops
    try:
        path = build("layer_ops.cu")
    except RuntimeError as ex:
        pytest.skip(f"emulator build unavailable: {ex}")
    lib = lops.use_emulated_library(path)
    try:
        yield lib
    finally:
        lops.use_emulated_library(None)
    assert lib.flpr_emu_deadlocks() == 0, "a kernel left threads parked at a barrier"


def test_launch_rewriter_handles_templates_casts_and_nested_calls():
    from emu.build_emu import rewrite_launches
    src = ("  k1<<<grid_for(n / 4, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(a), f(b, c), n / 4);\n"
           "  if (x) ns::k2<8><<<(unsigned)rows, 128>>>(p, q);\n")
    out, n = rewrite_launches(src)
    assert n == 2 and "<<<" not in out
    assert "flpr_emu::launch(flpr_emu::dims(grid_for(n / 4, 256)), flpr_emu::dims(256), (size_t)(0), (const void*)(st), " \
           "[=]() { k1(reinterpret_cast<const float4*>(a), f(b, c), n / 4); });" in out
    assert "if (x) flpr_emu::launch(flpr_emu::dims((unsigned)rows), flpr_emu::dims(128), 0, nullptr, " \
           "[=]() { ns::k2<8>(p, q); });" in out


def test_source_rewrites_for_shared_memory_and_inline_ptx():
    from emu.build_emu import rewrite_shared, strip_inline_ptx_functions
    out, n = rewrite_shared("  __shared__ float sh[2][8][256];\n  __shared__ float s_inv;\n  extern __shared__ int dyn[];\n")
    assert n == 2 and "__shared__" not in out
    assert "int* dyn = reinterpret_cast<int*>(flpr_emu::dyn_shared());" in out
    assert "using flpr_sh_t1 = float[2][8][256]; flpr_sh_t1& sh = *flpr_emu::shared<flpr_sh_t1>(1);" in out
    assert "using flpr_sh_t2 = float; flpr_sh_t2& s_inv = *flpr_emu::shared<flpr_sh_t2>(2);" in out
    src = ("int keep() { return 1; }\n"
           "__device__ __forceinline__ unsigned long long gtimer() {\n  unsigned long long t;\n"
           "  asm volatile(\"mov.u64 %0, %globaltimer;\" : \"=l\"(t));\n  return t;\n}\n"
           "int keep2() { return 2; }\n")
    out, removed = strip_inline_ptx_functions(src)
    assert removed == ["gtimer"] and "asm" not in out and "keep()" in out and "keep2()" in out


def test_every_self_check_family_passes_on_the_emulated_kernels(emu):
    """What ``ops.layer.enabled`` runs on the device before a family's first use - here on the emulated kernels."""
    from flpr_b200.ops import layer as lops, native
    before = native.launches()
    verdict = lops.run_checks_inprocess("cpu")
    assert verdict == {"wcompose": True, "swin_tokens": True, "apply": True, "ln_train": True}, verdict
    assert native.launches() - before >= 50, "the checks did not take the kernel path"
    assert emu.flpr_emu_deadlocks() == 0


@pytest.mark.parametrize("channels_last", [False, True])
def test_compose_kernels_emulated(emu, channels_last):
    L.check_compose_function("cpu", channels_last, use_ref=False)


@pytest.mark.parametrize("channels_last", [False, True])
def test_fedweit_layer_fused_theta_emulated(emu, channels_last):
    L.check_fedweit_layer_fused("cpu", channels_last, use_ref=False)


def test_atten_composer_emulated(emu):
    L.check_atten_composer_storage("cpu", use_ref=False)


def test_apply_global_kernel_emulated(emu):
    L.check_apply_global("cpu")


def test_swin_token_kernels_emulated(emu):
    L.check_swin_token_ops("cpu")


def test_gelu_act_emulated(emu):
    L.check_gelu_act("cpu")


def test_window_merge_residual_emulated(emu):
    L.check_window_merge_residual("cpu")


def test_layer_norm_rows_emulated(emu):
    L.check_layer_norm_rows("cpu")


def test_emulator_reports_a_barrier_not_every_thread_reaches(tmp_path):
    """The emulator's own failure mode: a full-mask shuffle under divergent control flow is reported, not hung on."""
    import ctypes
    import os
    import subprocess
    from emu.build_emu import CUDA_INCLUDE, HERE
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    src = tmp_path / "bad.cpp"
    src.write_text('#include "cuda_emu.h"\n'
                   "__global__ void bad_kernel(int* out) {\n"
                   "  int v = (int)threadIdx.x;\n"                         # a full-mask shuffle only half of each warp
                   "  if ((threadIdx.x & 31) < 16) v = __shfl_xor_sync(0xffffffffu, v, 1);\n"       # reaches ...
                   "  __syncthreads();\n"                                  # ... while the other half waits here
                   "  out[threadIdx.x] = v;\n"
                   "}\n"
                   "__global__ void good_kernel(float* out) {\n"
                   "  float v = (float)threadIdx.x;\n"
                   "  v = flpr::warp_sum(v);\n"
                   "  __syncthreads();\n"
                   "  out[threadIdx.x] = v;\n"
                   "}\n"
                   'extern "C" void run_bad(int* out) { flpr_emu::launch(flpr_emu::dims(1), flpr_emu::dims(64), 0, nullptr, [=]() { bad_kernel(out); }); }\n'
                   'extern "C" void run_good(float* out) { flpr_emu::launch(flpr_emu::dims(1), flpr_emu::dims(64), 0, nullptr, [=]() { good_kernel(out); }); }\n')
    lib_path = str(tmp_path / "libbad.so")
    subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-Wno-attributes", "-I", HERE, "-I", CUDA_INCLUDE,
                    str(src), "-o", lib_path], check=True)
    lib = ctypes.CDLL(lib_path)
    out = torch.zeros(64)
    lib.run_good(ctypes.c_void_p(out.data_ptr()))
    assert out[:32].eq(sum(range(32))).all() and out[32:].eq(sum(range(32, 64))).all()
    assert lib.flpr_emu_deadlocks() == 0
    bad = torch.zeros(64, dtype=torch.int32)
    lib.run_bad(ctypes.c_void_p(bad.data_ptr()))
    assert lib.flpr_emu_deadlocks() == 1
    assert os.path.exists(lib_path)


def test_emulator_primitives(tmp_path):
    """The emulator's own building blocks on a purpose-written kernel: 2-D grids, multi-dimensional blocks, dynamic and
    static shared memory (per block, zero-initialised), ``atomicAdd``, ``__syncthreads_or``, ``__shfl_down_sync``."""
    import ctypes
    import subprocess
    from emu.build_emu import CUDA_INCLUDE, HERE, rewrite_launches, rewrite_shared
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    cu = ('#include "cuda_emu.h"\n'
          "__global__ void k(int* out, int* votes, float* red, int n) {\n"
          "  extern __shared__ int dyn[];\n"
          "  __shared__ int s_count, s_flag[2];\n"
          "  const int lin = threadIdx.y * blockDim.x + threadIdx.x;\n"
          "  const int b = blockIdx.y * gridDim.x + blockIdx.x;\n"
          "  dyn[lin] = lin + b;\n"
          "  atomicAdd(&s_count, 1);\n"
          "  const int any = __syncthreads_or(lin == 5 && b == 1);\n"
          "  if (lin == 0) { out[b] = s_count + dyn[blockDim.x * blockDim.y - 1]; votes[b] = any; s_flag[0] = b; }\n"
          "  float v = (float)(lin & 31);\n"
          "  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);\n"
          "  if ((lin & 31) == 0) red[b * (n / 32) + lin / 32] = v;\n"
          "}\n"
          'extern "C" void run(int* out, int* votes, float* red) {\n'
          "  k<<<dim3(2, 3), dim3(16, 4), 64 * sizeof(int), 0>>>(out, votes, red, 64);\n"
          "}\n")
    src, n_sh = rewrite_shared(cu)
    src, n_l = rewrite_launches(src)
    assert n_sh == 2 and n_l == 1
    path = tmp_path / "prim.cpp"
    path.write_text(src)
    lib_path = str(tmp_path / "libprim.so")
    subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-Wno-attributes", "-Wno-unused-function", "-I", HERE,
                    "-I", CUDA_INCLUDE, str(path), "-o", lib_path], check=True)
    lib = ctypes.CDLL(lib_path)
    out, votes = torch.zeros(6, dtype=torch.int32), torch.full((6,), -1, dtype=torch.int32)
    red = torch.zeros(6 * 2)
    lib.run(ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(votes.data_ptr()), ctypes.c_void_p(red.data_ptr()))
    assert out.tolist() == [64 + 63 + b for b in range(6)]              # 64 threads counted, dyn[63] = 63 + b
    assert votes.tolist() == [0, 1, 0, 0, 0, 0]                          # only block 1 had a voting thread
    assert torch.equal(red, torch.full((12,), float(sum(range(32)))))   # lane 0 holds the warp sum
    assert lib.flpr_emu_deadlocks() == 0
