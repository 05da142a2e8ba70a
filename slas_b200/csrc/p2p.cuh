// p2p.cuh -- the fused collective of the sharded reductions: instead of "kernel, then ncclAllReduce,
// then finish kernel", the LAST block of the reduction kernel exchanges the rank-local f64 partial
// with every peer GPU directly over NVLink peer memory (st.global on IPC-mapped mailboxes), waits for
// the peers' partials and adds them in RANK ORDER -- so every rank ends with the same bits, in the same
// launch that read the data.  One process per GPU; the kernels of the N ranks run on N different GPUs
// at the same time (the ranks call the sharded ops collectively, like any collective).
#pragma once
#include <stdint.h>

namespace slas {

constexpr int kMaxPeers = 8;

struct PeerSlot {            // 32 bytes
    double v[2];
    unsigned long long flag; // epoch of the value
    unsigned long long pad;
};

struct PeerExchange {
    PeerSlot *mailbox[kMaxPeers];  // mailbox[r]: rank r's mailbox as mapped into THIS rank (own included)
    int nranks = 1, rank = 0;
    unsigned long long epoch = 0;  // strictly increasing, identical on all ranks for the same collective
};
// A mailbox holds 2 (epoch parity) x kMaxPeers slots.  Parity double-buffering is enough: a rank can
// only be one collective ahead of the slowest peer, because finishing epoch e + 1 needs that peer's
// epoch e + 1 value, which it sends only after it has read everything of epoch e.
constexpr size_t kMailboxBytes = 2 * kMaxPeers * sizeof(PeerSlot);

#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long ld_flag_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_flag_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t p2p_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Called by ALL threads of one block; p / p_im must be valid in threads [0, nranks).  On return thread 0
// holds the rank-ordered sums.  smem: 2 * kMaxPeers doubles.
__device__ __forceinline__ void peer_allreduce_sum(const PeerExchange &px, double &p, double &p_im, double *smem) {
    const int t = threadIdx.x;
    const unsigned parity = (unsigned)(px.epoch & 1ull);
    if (t < px.nranks) {
        PeerSlot *dst = px.mailbox[t] + parity * kMaxPeers + px.rank;  // my slot in rank t's mailbox
        volatile double *dv = dst->v;
        dv[0] = p;
        dv[1] = p_im;
        __threadfence_system();
        st_flag_sys(&dst->flag, px.epoch);
    }
    if (t < px.nranks) {
        PeerSlot *src = px.mailbox[px.rank] + parity * kMaxPeers + t;  // rank t's slot in my mailbox
        const uint64_t t0 = p2p_timer_ns();
        unsigned spins = 0;
        while (ld_flag_sys(&src->flag) != px.epoch) {
            // a peer that never arrives (crashed rank) must not hang the GPU forever
            if ((++spins & 0xfff) == 0 && p2p_timer_ns() - t0 > 20000000000ull) __trap();
        }
        const volatile double *sv = src->v;
        smem[2 * t] = sv[0];
        smem[2 * t + 1] = sv[1];
    }
    __syncthreads();
    if (t == 0) {
        double s = 0.0, s_im = 0.0;
        for (int r = 0; r < px.nranks; ++r) { s += smem[2 * r]; s_im += smem[2 * r + 1]; }
        p = s;
        p_im = s_im;
    }
}
#endif

}  // namespace slas
