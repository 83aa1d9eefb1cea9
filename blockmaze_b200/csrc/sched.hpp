// Which GPU takes the next proof.  Independent transaction proofs shard across the GPUs of one box (BASELINE.json north_star)
// with no data-path traffic between them, so the "scheduler" is a counter per device: a caller takes the device with the fewest proofs
// in flight (ties go round-robin, so an idle box fills device by device in rotation) and gives it back when its proof is done.
// Host-only and free of CUDA so that the policy is unit-tested on CPU (tests/test_host.py) through the zkb200_sched_* hooks.
#pragma once
#include <climits>
#include <cstdint>
#include <mutex>
#include <vector>

namespace zkb {

class DeviceSched {
  public:
    explicit DeviceSched(int n) : inflight_(n > 0 ? n : 1, 0), total_(n > 0 ? n : 1, 0) {}
    int size() const { return (int)inflight_.size(); }
    // slot of the least-loaded device; the caller owns one unit of its load until done(slot)
    int pick() {
        std::lock_guard<std::mutex> lk(mu_);
        const int n = (int)inflight_.size();
        int best = 0, load = INT_MAX;
        for (int k = 0; k < n; k++) {
            const int i = (next_ + k) % n;
            if (inflight_[i] < load) { load = inflight_[i]; best = i; }
        }
        inflight_[best]++; total_[best]++;
        next_ = (best + 1) % n;
        return best;
    }
    void done(int slot) {
        std::lock_guard<std::mutex> lk(mu_);
        if (slot >= 0 && slot < (int)inflight_.size() && inflight_[slot] > 0) inflight_[slot]--;
    }
    int inflight(int slot) { std::lock_guard<std::mutex> lk(mu_); return slot >= 0 && slot < (int)inflight_.size() ? inflight_[slot] : -1; }
    uint64_t total(int slot) { std::lock_guard<std::mutex> lk(mu_); return slot >= 0 && slot < (int)total_.size() ? total_[slot] : 0; }

  private:
    std::mutex mu_;
    std::vector<int> inflight_;
    std::vector<uint64_t> total_;
    int next_ = 0;
};

} // namespace zkb
