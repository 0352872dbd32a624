// ptx_sim.h -- TEST INFRASTRUCTURE: what the inline-PTX helpers of partition.cuh do, for the CPU stand-in.
// An mbarrier (count 1) is a 64-bit word counting completed phases; a TMA bulk load is a memcpy by the issuing thread
// followed by the completion of the phase; wait(parity) spins until the phase of that parity has completed.
#pragma once
#include <sched.h>
static inline uint32_t ok_smem_u32(const void* p) { return (uint32_t)(uintptr_t)p; }
static inline void ok_mbar_init(unsigned long long* bar, unsigned) { __atomic_store_n(bar, 0ull, __ATOMIC_SEQ_CST); }
static inline void ok_mbar_arrive(unsigned long long* bar) { __atomic_fetch_add(bar, 1ull, __ATOMIC_SEQ_CST); }
static inline void ok_tma_load_1d(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    memcpy(dst_smem, src_gmem, bytes);
    __atomic_fetch_add(bar, 1ull, __ATOMIC_SEQ_CST);
}
static inline void ok_mbar_wait(unsigned long long* bar, unsigned phase) {
    while ((__atomic_load_n(bar, __ATOMIC_SEQ_CST) & 1ull) == (unsigned long long)(phase & 1u)) sched_yield();
}
static inline void ok_tma_store_1d(void* dst_gmem, const void* src_smem, unsigned bytes) { memcpy(dst_gmem, src_smem, bytes); }
static inline void ok_tma_store_wait_read() {}
static inline void ok_tma_store_wait_all() {}
