// multi.cuh -- several GPUs inside ONE process (multi.cu).  A slas backend is a zero-sized type made by
// `B::default()` in an ordinary program (src/static_vec.rs:76,219, src/backends/rust.rs:2-3): no launcher, no ranks.
// After slas_cuda_init_devices(n) (or SLAS_CUDA_DEVICES=n in the environment):
//   * an op on DEVICE operands runs on the device that owns them (DevScope);
//   * unit-parallel ops on HOST operands (element-wise, every *_batched entry) are split over the devices by element
//     range / batch index, each device running its own 3-slot copy pipeline on its own worker thread (multi_run);
//   * dot / norm / normalize on HOST operands are split by vector slice and the f64 partials are added in device
//     order on the host;
//   * the *_multi entries take one device-resident slice per device and do the exchange inside the reduction kernels
//     over NVLink peer memory -- the in-process form of the *_sharded ops (no IPC: peer access).
#pragma once
#include <condition_variable>
#include <functional>

#include "common.cuh"
#include "p2p.cuh"

namespace slas {

int multi_ndev();        // 1 unless slas_cuda_init_devices initialised more
void multi_autoinit();   // honours SLAS_CUDA_DEVICES once per process

// For the scope: the current device is the one that owns `p` when p is a device pointer and several devices are
// initialised; otherwise nothing changes.
struct DevScope {
    int prev = -1;
    explicit DevScope(const void *p);
    ~DevScope();
    DevScope(const DevScope &) = delete;
    DevScope &operator=(const DevScope &) = delete;
};

// Several devices, `p` a host pointer, and enough bytes for a split to pay for itself.
bool multi_split_wanted(const void *p, size_t bytes);

// fn(d) runs on device d's worker thread (device d current) for d in [0, parts), all at once; the first failure
// (status and message) is returned to the caller.  One split op at a time per process.
int multi_run(int parts, const std::function<int(int)> &fn);

// Meeting point for the `parts` tasks of one multi_run (sum the partials, then go on): every task calls arrive()
// exactly once, failed or not; arrive returns false if any task failed before it.
struct Rendezvous {
    explicit Rendezvous(int parts) : waiting(parts) {}
    bool arrive(bool ok) {
        std::unique_lock<std::mutex> lk(mu);
        if (!ok) all_ok = false;
        if (--waiting == 0) cv.notify_all();
        else cv.wait(lk, [&] { return waiting == 0; });
        return all_ok;
    }
    std::mutex mu;
    std::condition_variable cv;
    int waiting;
    bool all_ok = true;
};

// Peer-memory exchange descriptor for slice `rank` of `nslices`, slice i living on devices[i]; false when the mailboxes
// are not available (no peer access between the devices): the caller combines partials on the host instead.
bool multi_peer_exchange(const int *devices, int nslices, int rank, unsigned long long epoch, PeerExchange *out);
unsigned long long multi_next_epoch();
Context *multi_context(int device);  // initialised context of `device`, or nullptr

}  // namespace slas
