// helper_pool.hpp — parked threads for the host-side copies of a call.  No CUDA: tests/test_helper_pool.py compiles
// this header into a small harness and stresses it on the CPU.
#pragma once

#include <algorithm>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace acb200 {

// A few parked threads per handle and device slot for the host-side copies of a call (the gather of pageable or
// scattered haystacks into pinned staging).  Creating a dozen threads per slab or per call costs 0.1-0.2 ms — more
// than the copy of a mid-size haystack they are there to speed up.  run() hands out the indices 0..n-1 one at a time
// (the caller takes part) and returns when all of them are done.
class HelperPool {
public:
    explicit HelperPool(int n_threads) : wanted_(n_threads), limit_(n_threads)
    {
        try {
            for (int i = 0; i < n_threads; ++i) th_.emplace_back([this, i] { loop(i); });
        } catch (...) {}        // (thread limit reached: fewer helpers, the caller of run() does the rest itself)
    }
    ~HelperPool()
    {
        { std::lock_guard<std::mutex> g(m_); quit_ = true; }
        cv_.notify_all();
        for (auto &x : th_) x.join();
    }
    int helpers() const { return std::min(limit_, (int)th_.size()); }     // helpers that take part in run()
    int wanted() const { return wanted_; }
    // a call that shares the host's cores with other device slots uses only the first `limit` helpers (between calls only)
    void set_limit(int limit) { std::lock_guard<std::mutex> g(m_); limit_ = std::max(0, limit); }
    void run(int n, const std::function<void(int)> &f)
    {
        if (n <= 0) return;
        std::unique_lock<std::mutex> g(m_);
        job_ = &f; n_ = n; next_ = 0; pending_ = n;
        g.unlock();
        if (n > 1) cv_.notify_all();
        g.lock();
        while (next_ < n_) {
            const int i = next_++;
            g.unlock();
            f(i);
            g.lock();
            --pending_;
        }
        done_.wait(g, [this] { return pending_ == 0; });
        job_ = nullptr;
    }
private:
    void loop(int id)
    {
        std::unique_lock<std::mutex> g(m_);
        while (true) {
            cv_.wait(g, [this, id] { return quit_ || (job_ && next_ < n_ && id < limit_); });
            if (quit_) return;
            const std::function<void(int)> *f = job_;
            const int i = next_++;
            g.unlock();
            (*f)(i);
            g.lock();
            if (--pending_ == 0) done_.notify_all();
        }
    }
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(int)> *job_ = nullptr;
    int n_ = 0, next_ = 0, pending_ = 0;
    bool quit_ = false;
    int wanted_ = 0, limit_ = 0;
};

} // namespace acb200
