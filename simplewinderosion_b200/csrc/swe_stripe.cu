// swe_stripe.cu -- multi-GPU row stripes: halo push, particle hand-off and neighbour
// synchronisation done by kernels over peer-mapped memory (NVLink/NVSwitch P2P through CUDA IPC),
// with no host involvement inside a World::erode call.
//
// Decomposition: rank r owns the global rows [own_lo, own_hi) and stores
// [own_lo-5, own_hi+7) (clipped to the map). Per round of the parallel pipeline:
//   1. par_round_kernel<STRIPE>: steps the local particles (they only write their own, owned,
//      centre cell); a particle whose next cell belongs to a neighbour is stored straight into
//      that neighbour's inbox (peer stores + one system-scope atomic for the slot).
//   2. wait until the neighbours have finished READING the ghosts of the previous round
//      (consumed flags), then push the post-deposit boundary rows (heights + stamps) into their
//      ghost rows: my first 7 owned rows go below, my last 5 owned rows go above.
//   3. signal "round r delivered" (ready flags), wait for the neighbours' ready flags.
//   4. merge the inbox into the local particle arrays.
//   5. stamped cascade on rows [own_lo-1, own_hi+1): with the 5/7-row halo the nine-colour
//      dependency cone (up to 4 rows low, 6 high, depending on alignment) is complete, so these rows come out bitwise equal to
//      what the neighbour computes for them -- no second exchange for the next round's reads.
//   6. signal "ghosts of round r consumed".
// Flags are monotonically increasing sequence numbers written with system-scope release stores
// after a __threadfence_system() and polled with volatile loads; every rank posts before it waits,
// so there is no cyclic wait. Polls give up after ~10 s and raise an error flag instead of
// hanging the GPU.
#include "swe_common.cuh"

namespace swe {

__global__ void stripe_push_rows_kernel(const float* __restrict__ src_h, float* __restrict__ dst_h,
                                        const unsigned char* __restrict__ src_s,
                                        unsigned char* __restrict__ dst_s, size_t cells) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells;
       i += (size_t)gridDim.x * blockDim.x) {
    dst_h[i] = src_h[i];
    if (src_s) dst_s[i] = src_s[i];
  }
}

__global__ void stripe_signal_kernel(unsigned* peer_flag_a, unsigned* peer_flag_b, unsigned value) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  __threadfence_system();   // everything this stream wrote to peer memory is visible first
  if (peer_flag_a) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer_flag_a), "r"(value) : "memory");
  if (peer_flag_b) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer_flag_b), "r"(value) : "memory");
}

__global__ void stripe_wait_kernel(const unsigned* flag_a, const unsigned* flag_b, unsigned value,
                                   unsigned* error_flag) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (*reinterpret_cast<volatile unsigned*>(error_flag)) return;   // already failed: do not wait again
  const long long t0 = clock64();
  const unsigned* flags[2] = {flag_a, flag_b};
  for (int k = 0; k < 2; k++) {
    if (!flags[k]) continue;
    for (;;) {
      unsigned v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags[k]) : "memory");
      if ((int)(v - value) >= 0) break;
      if (clock64() - t0 > 20000000000LL) { atomicExch(error_flag, 1u); return; }
      __nanosleep(200);
    }
  }
  __threadfence_system();
}

// append the particles neighbours handed over in this round; the inbox counter is reset for its
// next use (two rounds later, after this rank has signalled again)
__global__ void stripe_merge_kernel(const float4* __restrict__ in_a, const float4* __restrict__ in_b,
                                    const uint32_t* __restrict__ in_pid, unsigned* in_count, int in_capacity,
                                    float4* __restrict__ pa, float4* __restrict__ pb,
                                    uint32_t* __restrict__ pid, unsigned* n_used, int capacity,
                                    unsigned* overflow, unsigned* base_slot) {
  const unsigned cnt = min(*in_count, (unsigned)in_capacity);
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cnt) {
    const unsigned slot = *base_slot + i;
    if (slot < (unsigned)capacity) {
      pa[slot] = in_a[i]; pb[slot] = in_b[i]; pid[slot] = in_pid[i];
    } else {
      atomicAdd(overflow, 1u);
    }
  }
}
__global__ void stripe_merge_prepare_kernel(const unsigned* in_count, int in_capacity, const unsigned* n_used,
                                            unsigned* base_slot, unsigned* overflow) {
  *base_slot = *n_used;
  if (*in_count > (unsigned)in_capacity) atomicAdd(overflow, *in_count - (unsigned)in_capacity);
}
__global__ void stripe_merge_finish_kernel(unsigned* in_count, int in_capacity, unsigned* n_used, int capacity) {
  const unsigned cnt = min(*in_count, (unsigned)in_capacity);
  const unsigned nu = *n_used + cnt;
  *n_used = nu < (unsigned)capacity ? nu : (unsigned)capacity;
  *in_count = 0u;
}

void launch_stripe_push(const float* src_h, float* dst_h, const unsigned char* src_s, unsigned char* dst_s,
                        size_t cells, cudaStream_t st) {
  if (cells == 0) return;
  int blocks = (int)((cells + 255) / 256);
  if (blocks > 148 * 4) blocks = 148 * 4;
  stripe_push_rows_kernel<<<blocks, 256, 0, st>>>(src_h, dst_h, src_s, dst_s, cells);
}
void launch_stripe_signal(unsigned* peer_flag_a, unsigned* peer_flag_b, unsigned value, cudaStream_t st) {
  if (!peer_flag_a && !peer_flag_b) return;
  stripe_signal_kernel<<<1, 32, 0, st>>>(peer_flag_a, peer_flag_b, value);
}
void launch_stripe_wait(const unsigned* flag_a, const unsigned* flag_b, unsigned value, unsigned* error_flag,
                        cudaStream_t st) {
  if (!flag_a && !flag_b) return;
  stripe_wait_kernel<<<1, 32, 0, st>>>(flag_a, flag_b, value, error_flag);
}
void launch_stripe_merge(const float4* in_a, const float4* in_b, const uint32_t* in_pid, unsigned* in_count,
                         int in_capacity, float4* pa, float4* pb, uint32_t* pid, unsigned* n_used,
                         int capacity, unsigned* overflow, unsigned* scratch_base, cudaStream_t st) {
  stripe_merge_prepare_kernel<<<1, 1, 0, st>>>(in_count, in_capacity, n_used, scratch_base, overflow);
  stripe_merge_kernel<<<(in_capacity + 255) / 256, 256, 0, st>>>(in_a, in_b, in_pid, in_count, in_capacity, pa,
                                                                  pb, pid, n_used, capacity, overflow, scratch_base);
  stripe_merge_finish_kernel<<<1, 1, 0, st>>>(in_count, in_capacity, n_used, capacity);
}

// ---- fused exchange kernels (one launch each instead of 4 + 7) -----------------------------------
// poll helper: thread 0 of the calling block spins until both flags reached `value`
__device__ __forceinline__ bool spin_until(const unsigned* flag_a, const unsigned* flag_b, unsigned value,
                                           unsigned* error_flag) {
  if (*reinterpret_cast<volatile unsigned*>(error_flag)) return false;
  const long long t0 = clock64();
  const unsigned* flags[2] = {flag_a, flag_b};
  for (int k = 0; k < 2; k++) {
    if (!flags[k]) continue;
    for (;;) {
      unsigned v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags[k]) : "memory");
      if ((int)(v - value) >= 0) break;
      if (clock64() - t0 > 20000000000LL) { atomicExch(error_flag, 1u); return false; }
      __nanosleep(100);
    }
  }
  __threadfence_system();
  return true;
}

// wait(consumed >= seq-1) -> push both boundary strips into the neighbours' ghost rows -> the last
// block to finish signals ready = seq to both neighbours.
__global__ void __launch_bounds__(256)
stripe_push_signal_kernel(StripePush down, StripePush up, const unsigned* consumed_a, const unsigned* consumed_b,
                          unsigned* ready_a, unsigned* ready_b, unsigned seq, unsigned* done_counter,
                          unsigned* error_flag) {
  __shared__ int ok;
  if (threadIdx.x == 0) ok = spin_until(consumed_a, consumed_b, seq - 1u, error_flag) ? 1 : 0;
  __syncthreads();
  if (ok) {
    const size_t stride = (size_t)gridDim.x * blockDim.x, t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (size_t i = t; i < down.cells; i += stride) {
      down.dst_h[i] = down.src_h[i];
      if (down.src_s) down.dst_s[i] = down.src_s[i];
    }
    for (size_t i = t; i < up.cells; i += stride) {
      up.dst_h[i] = up.src_h[i];
      if (up.src_s) up.dst_s[i] = up.src_s[i];
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(done_counter, 1u);
    if (prev == gridDim.x - 1) {               // every block has pushed (and fenced) its share
      *done_counter = 0u;
      __threadfence_system();
      if (ready_a) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(ready_a), "r"(seq) : "memory");
      if (ready_b) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(ready_b), "r"(seq) : "memory");
    }
  }
}

// wait(ready >= seq) -> append both inboxes to the local particle arrays -> the last block bumps
// n_used and clears the inbox counters.
__global__ void __launch_bounds__(256)
stripe_wait_merge_kernel(StripeOutbox in_a, StripeOutbox in_b, const unsigned* ready_a, const unsigned* ready_b,
                         unsigned seq, float4* __restrict__ pa, float4* __restrict__ pb,
                         uint32_t* __restrict__ pid, unsigned* n_used, int capacity, unsigned* overflow,
                         unsigned* done_counter, unsigned* error_flag) {
  __shared__ int ok;
  if (threadIdx.x == 0) ok = spin_until(ready_a, ready_b, seq, error_flag) ? 1 : 0;
  __syncthreads();
  const unsigned base = *n_used;               // constant until the last block updates it
  unsigned ca = 0, cb = 0;
  if (ok) {
    ca = in_a.count ? min(*in_a.count, (unsigned)in_a.capacity) : 0u;
    cb = in_b.count ? min(*in_b.count, (unsigned)in_b.capacity) : 0u;
    const unsigned stride = gridDim.x * blockDim.x, t = blockIdx.x * blockDim.x + threadIdx.x;
    for (unsigned i = t; i < ca + cb; i += stride) {
      const unsigned slot = base + i;
      if (slot >= (unsigned)capacity) { atomicAdd(overflow, 1u); continue; }
      if (i < ca) { pa[slot] = in_a.a[i]; pb[slot] = in_a.b[i]; pid[slot] = in_a.pid[i]; }
      else { const unsigned j = i - ca; pa[slot] = in_b.a[j]; pb[slot] = in_b.b[j]; pid[slot] = in_b.pid[j]; }
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(done_counter, 1u);
    if (prev == gridDim.x - 1) {
      *done_counter = 0u;
      if (ok) {
        if (in_a.count && *in_a.count > (unsigned)in_a.capacity) atomicAdd(overflow, *in_a.count - (unsigned)in_a.capacity);
        if (in_b.count && *in_b.count > (unsigned)in_b.capacity) atomicAdd(overflow, *in_b.count - (unsigned)in_b.capacity);
        const unsigned nu = base + ca + cb;
        *n_used = nu < (unsigned)capacity ? nu : (unsigned)capacity;
        if (in_a.count) *in_a.count = 0u;
        if (in_b.count) *in_b.count = 0u;
      }
    }
  }
}

void launch_stripe_push_signal(const StripePush& down, const StripePush& up, const unsigned* consumed_a,
                               const unsigned* consumed_b, unsigned* ready_a, unsigned* ready_b, unsigned seq,
                               unsigned* done_counter, unsigned* error_flag, cudaStream_t st) {
  stripe_push_signal_kernel<<<64, 256, 0, st>>>(down, up, consumed_a, consumed_b, ready_a, ready_b, seq,
                                                done_counter, error_flag);
}
void launch_stripe_wait_merge(const StripeOutbox& in_a, const StripeOutbox& in_b, const unsigned* ready_a,
                              const unsigned* ready_b, unsigned seq, float4* pa, float4* pb, uint32_t* pid,
                              unsigned* n_used, int capacity, unsigned* overflow, unsigned* done_counter,
                              unsigned* error_flag, cudaStream_t st) {
  stripe_wait_merge_kernel<<<32, 256, 0, st>>>(in_a, in_b, ready_a, ready_b, seq, pa, pb, pid, n_used, capacity,
                                               overflow, done_counter, error_flag);
}

// Force-load this translation unit's kernels (CUDA loads modules lazily on first launch, and a
// first launch issued while another stream spins in a stripe wait kernel would stall behind it).
void preload_stripe() {
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, stripe_push_rows_kernel);
  cudaFuncGetAttributes(&a, stripe_signal_kernel);
  cudaFuncGetAttributes(&a, stripe_wait_kernel);
  cudaFuncGetAttributes(&a, stripe_merge_kernel);
  cudaFuncGetAttributes(&a, stripe_merge_prepare_kernel);
  cudaFuncGetAttributes(&a, stripe_merge_finish_kernel);
  cudaFuncGetAttributes(&a, stripe_push_signal_kernel);
  cudaFuncGetAttributes(&a, stripe_wait_merge_kernel);
}

}  // namespace swe
