// Shared helpers for libsfa_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "sfa_b200.h"

namespace sfa {

// thread-local last-error message (sfa_last_error)
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define SFA_CUDA_TRY(expr)                                         \
    do {                                                           \
        cudaError_t _e = (expr);                                   \
        if (_e != cudaSuccess) return ::sfa::cuda_fail(_e, #expr); \
    } while (0)

#define SFA_REQUIRE(cond, ...)                 \
    do {                                       \
        if (!(cond)) {                         \
            ::sfa::set_error(__VA_ARGS__);     \
            return SFA_ERR_INVALID_ARGUMENT;   \
        }                                      \
    } while (0)

// ---- launch accounting / optional per-kernel event timing (host_pipeline.cu) ----
void note_launch();
bool profiling_on();
void profile_mark(const char* name, cudaStream_t stream, bool begin);

// Wraps one kernel launch statement: counts it and, while profiling, brackets it with events.
#define SFA_LAUNCH(name, stream, ...)                                       \
    do {                                                                    \
        const bool _prof = ::sfa::profiling_on();                           \
        if (_prof) ::sfa::profile_mark(name, stream, true);                 \
        __VA_ARGS__;                                                        \
        ::sfa::note_launch();                                               \
        if (_prof) ::sfa::profile_mark(name, stream, false);                \
    } while (0)

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// ---- L2 eviction policy for write-once / read-once streams ---------------------------------------
// The sweeps (read once) and the BEV planes (written once) stream THROUGH the L2, while the point
// records in between are written by bev_bin and read back by bev_band microseconds later: marking
// the streams evict-first keeps them from pushing the records out to HBM.
__device__ __forceinline__ unsigned long long l2_evict_first_policy() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

// ---- streaming memory access (data a CTA touches exactly once: keep it out of L1) ----
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ld_stream_f4_evict_first(const float4* p, unsigned long long pol) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void st_stream_f4(float4* p, const float4& v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

// ---- shared-memory access through a 32-bit shared::cta address ----------------------------------
// For sm_100a nvcc rebuilds the shared window base (S2UR SR_CgaCtaId + three uniform ops) at every access to a __shared__
// array; in an issue-bound kernel those are a tenth of the instructions.  The hot sites take ONE base register
// (smem_addr_u32 of the kernel's shared struct) plus compile-time offsets instead.
__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory");
    return v;
}
// if (pred) st.shared.v4
__device__ __forceinline__ void sts_v4_if(bool pred, uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("{ .reg .pred p; setp.ne.u32 p, %5, 0; @p st.shared.v4.u32 [%0], {%1,%2,%3,%4}; }" ::"r"(saddr), "r"(a), "r"(b),
                 "r"(c), "r"(d), "r"((uint32_t)pred)
                 : "memory");
}
// pred ? atomicAdd(shared u32, 1) : `otherwise`
__device__ __forceinline__ uint32_t atoms_inc_if(bool pred, uint32_t saddr, uint32_t otherwise) {
    uint32_t old;
    asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; mov.u32 %0, %3; @p atom.shared.add.u32 %0, [%1], 1; }"
                 : "=r"(old)
                 : "r"(saddr), "r"((uint32_t)pred), "r"(otherwise)
                 : "memory");
    return old;
}

// the same with a compile-time byte offset folded into the address ([reg + imm])
template <int OFF> __device__ __forceinline__ uint32_t lds_at(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1 + %2];" : "=r"(v) : "r"(saddr), "n"(OFF) : "memory");
    return v;
}
template <int OFF> __device__ __forceinline__ void sts_at(uint32_t saddr, uint32_t v) {
    asm volatile("st.shared.u32 [%0 + %1], %2;" ::"r"(saddr), "n"(OFF), "r"(v) : "memory");
}
template <int OFF> __device__ __forceinline__ void red_max_at(uint32_t saddr, uint32_t v) {
    asm volatile("red.shared.max.u32 [%0 + %1], %2;" ::"r"(saddr), "n"(OFF), "r"(v) : "memory");
}
template <int OFF> __device__ __forceinline__ void red_inc_at(uint32_t saddr) {
    asm volatile("red.shared.add.u32 [%0 + %1], 1;" ::"r"(saddr), "n"(OFF) : "memory");
}
template <int OFF> __device__ __forceinline__ uint32_t cas_at(uint32_t saddr, uint32_t expect, uint32_t desired) {
    uint32_t old;
    asm volatile("atom.shared.cas.b32 %0, [%1 + %2], %3, %4;" : "=r"(old) : "r"(saddr), "n"(OFF), "r"(expect), "r"(desired) : "memory");
    return old;
}

// ---- programmatic dependent launch (griddepcontrol) ----------------------------------------------
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may become resident once every CTA of the kernel
// in front of it has called grid_launch_dependents() (or exited); it must call grid_dependency_wait() before it touches anything
// that kernel wrote.  Both are no-ops in an ordinary launch.
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- TMA bulk copies (cp.async.bulk, 1-D, no tensor map) ----------------------------------------
__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
// Generic-proxy writes to shared memory (st.shared, atom.shared) -> visible to the async proxy.
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// shared::cta -> global, `bytes` a multiple of 16, both addresses 16-B aligned; joins the thread's bulk group.
__device__ __forceinline__ void bulk_store_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_addr_u32(ssrc)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_store_s2g_hint(void* gdst, const void* ssrc, uint32_t bytes, unsigned long long pol) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst),
                 "r"(smem_addr_u32(ssrc)), "r"(bytes), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all committed bulk groups of this thread have finished READING their shared-memory source
__device__ __forceinline__ void bulk_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... have completed entirely
__device__ __forceinline__ void bulk_wait_group0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- mbarriers (shared::cta) and global -> shared bulk copies completing on them ----------------
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_addr_u32(bar)), "r"(parity), "r"(200u)
        : "memory");
}
// the service thread's wait: the hardware may suspend the thread for up to ~1 us per probe instead of spinning
__device__ __forceinline__ void mbar_wait_relaxed(unsigned long long* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAITR_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONER_%=;\n"
        "bra WAITR_%=;\n"
        "DONER_%=:\n"
        "}\n" ::"r"(smem_addr_u32(bar)), "r"(parity), "r"(1000u)
        : "memory");
}
// global -> shared::cta bulk copy, completion (bytes) on an mbarrier of this CTA
__device__ __forceinline__ void bulk_load_g2s(void* sdst, const void* gsrc, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr_u32(sdst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_addr_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_test(unsigned long long* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_addr_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}


// ---- order-preserving float -> uint32 maps -------------------------------------------------------
// Total order of the float VALUES (-0.0 == +0.0); `nan_key` is where NaN goes.
__device__ __forceinline__ uint32_t orderable_u32(float f, uint32_t nan_key) {
    uint32_t b = __float_as_uint(f);
    if (b == 0x80000000u) b = 0u;  // -0.0 compares equal to +0.0
    uint32_t k = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return (f != f) ? nan_key : k;
}
__device__ __forceinline__ float orderable_to_float(uint32_t k) {
    uint32_t b = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
    return __uint_as_float(b);
}

}  // namespace sfa
