// flpr host runtime helpers (C ABI, driven through ctypes like the kernels).
//
//   flpr_memcpy_d2h_async / flpr_memcpy2d_d2h_async   device -> page-locked host copies issued straight into
//       CUDA-registered *file mappings* (runtime/mapped_store.py): a checkpoint is a set of DMAs into the file's own
//       pages, no staging buffer, no pickling, no CPU copy. The pitched form writes the prototypes of FedSTIL's
//       exemplar memory into the array slots of the reference-schema pickle (methods/fedstil.py:841,846) in one call.
//   flpr_host_register / flpr_host_unregister          cudaHostRegister of a mapping (portable flag, every device).
#include <cuda_runtime.h>
#include <stddef.h>

extern "C" {

int flpr_memcpy_d2h_async(void* dst_host, const void* src_dev, size_t nbytes, cudaStream_t stream) {
  if (nbytes == 0) return 0;
  return (int)cudaMemcpyAsync(dst_host, src_dev, nbytes, cudaMemcpyDeviceToHost, stream);
}

int flpr_memcpy2d_d2h_async(void* dst_host, size_t dst_pitch, const void* src_dev, size_t src_pitch, size_t width,
                            size_t height, cudaStream_t stream) {
  if (width == 0 || height == 0) return 0;
  return (int)cudaMemcpy2DAsync(dst_host, dst_pitch, src_dev, src_pitch, width, height, cudaMemcpyDeviceToHost, stream);
}

int flpr_host_register(void* ptr, size_t nbytes) {
  return (int)cudaHostRegister(ptr, nbytes, cudaHostRegisterPortable);
}

int flpr_host_unregister(void* ptr) { return (int)cudaHostUnregister(ptr); }

// A CUDA stream of its own. torch.cuda.Stream() hands out streams from a pool of 32 per device round-robin, so two
// Stream objects can be the SAME hardware queue: a client thread's "own" stream may then coincide with the stream
// another thread is capturing a CUDA graph on, and that client's eager work (e.g. its augmentation's random draws)
// lands inside the capture. Client, capture, copy and communication streams are therefore created here and wrapped
// with torch.cuda.ExternalStream.
int flpr_stream_create(void** out, int priority) {
  cudaStream_t s = nullptr;
  cudaError_t e = cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, priority);
  *out = (void*)s;
  return (int)e;
}

int flpr_stream_destroy(void* s) { return (int)cudaStreamDestroy((cudaStream_t)s); }

}  // extern "C"
