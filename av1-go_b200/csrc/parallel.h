// Small process-wide worker pool for the host parser: tiles of a frame (and GOP segments of a file) are independent
// symbol streams, the only parallelism the sequential parse has.  parallel_for(n, fn) runs fn(0..n-1); the caller takes part,
// so nested use (a segment worker parsing its frame's tiles) always makes progress even when every pool thread is busy.
#pragma once
#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace av1r {

class WorkerPool {
public:
    static WorkerPool& get() {
        static WorkerPool p;
        return p;
    }
    int size() const { return (int)threads_.size(); }

    // Per-thread switch: a caller that is itself one of >= #cores concurrent workers (a file with many GOP segments) runs its
    // loops inline -- handing chunks to pool threads would only add waiting on time-sliced helpers.
    static bool& nested_enabled() {
        static thread_local bool v = true;
        return v;
    }

    void parallel_for(int n, const std::function<void(int)>& fn) {
        if (n <= 0) return;
        if (n == 1 || threads_.empty() || !nested_enabled()) {
            for (int i = 0; i < n; i++) fn(i);
            return;
        }
        auto job = std::make_shared<Job>();
        job->n = n;
        job->fn = &fn;
        {
            std::lock_guard<std::mutex> lk(m_);
            jobs_.push_back(job);
        }
        cv_.notify_all();
        run(*job);                       // the caller works too
        std::unique_lock<std::mutex> lk(job->m);
        job->cv.wait(lk, [&] { return job->done.load() == n; });
    }

private:
    struct Job {
        int n = 0;
        const std::function<void(int)>* fn = nullptr;
        std::atomic<int> next{0}, done{0};
        std::mutex m;
        std::condition_variable cv;
    };
    std::vector<std::thread> threads_;
    std::deque<std::shared_ptr<Job>> jobs_;
    std::mutex m_;
    std::condition_variable cv_;
    bool stop_ = false;

    WorkerPool() {
        int n = (int)std::thread::hardware_concurrency();
        if (n <= 0) n = 4;
        n = std::min(n, 64) - 1;         // callers are workers too
        for (int i = 0; i < n; i++) threads_.emplace_back([this] { loop(); });
    }
    ~WorkerPool() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : threads_) t.join();
    }
    static void run(Job& j) {
        while (true) {
            const int i = j.next.fetch_add(1);
            if (i >= j.n) return;
            (*j.fn)(i);
            if (j.done.fetch_add(1) + 1 == j.n) {
                std::lock_guard<std::mutex> lk(j.m);
                j.cv.notify_all();
            }
        }
    }
    void loop() {
        while (true) {
            std::shared_ptr<Job> job;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return stop_ || !jobs_.empty(); });
                if (stop_) return;
                job = jobs_.front();
                if (job->next.load() >= job->n) {   // fully handed out: retire it from the queue
                    jobs_.pop_front();
                    continue;
                }
            }
            run(*job);
        }
    }
};

}  // namespace av1r
