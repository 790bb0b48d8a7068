// pth_todo.cu — entry points declared in the header whose kernels are still
// being written.  They fail loudly (PTH_ENOSUP); nothing falls back to the CPU.
#include "pth_common.cuh"

#define PTH_TODO(name)                                  \
  pth_set_error(#name ": not implemented in this build"); \
  return PTH_ENOSUP;

extern "C" {
int pth_pack_transitions(pth_ctx*, const uint8_t*, const uint8_t*, const float*, const float*, const float*, int64_t, uint8_t*, void*) { PTH_TODO(pth_pack_transitions) }
int pth_pack_allgather_p2p(pth_ctx*, const uint8_t*, const uint8_t*, const float*, const float*, const float*, int64_t, uint8_t* const*, int32_t, int32_t, void*) { PTH_TODO(pth_pack_allgather_p2p) }
}
