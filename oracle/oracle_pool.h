// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// One persistent thread pool for the multi-threaded (timing) mode of the restatement, standing in for the reference's global
// rayon pool (src/square_wheel/host.rs:62-81): threads are created once and reused by every stage of every frame, like
// rayon's workers, instead of being spawned per call.
#pragma once
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace oracle
{

class ThreadPool
{
public:
	static ThreadPool& instance()
	{
		static ThreadPool pool;
		return pool;
	}

	// Runs fn(i) for i in [0, num_tasks) on up to num_threads threads (the caller is one of them); tasks are handed out
	// dynamically one at a time, in index order.
	void parallel_for(uint32_t num_tasks, int num_threads, const std::function<void(uint32_t)>& fn)
	{
		if (num_tasks == 0)
			return;
		if (num_threads <= 1 || num_tasks == 1)
		{
			for (uint32_t i = 0; i < num_tasks; ++i)
				fn(i);
			return;
		}
		std::unique_lock<std::mutex> call_lock(call_mutex_); // one parallel region at a time
		ensure_workers(num_threads - 1);
		{
			std::lock_guard<std::mutex> lk(mutex_);
			fn_ = &fn;
			num_tasks_ = num_tasks;
			next_.store(0);
			active_workers_ = std::min<int>(num_threads - 1, (int)workers_.size());
			pending_ = active_workers_;
			++generation_;
		}
		wake_.notify_all();
		run_tasks();
		std::unique_lock<std::mutex> lk(mutex_);
		done_.wait(lk, [&] { return pending_ == 0; });
		fn_ = nullptr;
	}

private:
	ThreadPool() = default;
	~ThreadPool()
	{
		{
			std::lock_guard<std::mutex> lk(mutex_);
			stop_ = true;
			++generation_;
		}
		wake_.notify_all();
		for (auto& t : workers_)
			t.join();
	}
	void ensure_workers(int n)
	{
		while ((int)workers_.size() < n)
		{
			const int id = (int)workers_.size();
			std::lock_guard<std::mutex> lk(mutex_);
			workers_.emplace_back([this, id, gen = generation_]() mutable { worker(id, gen); });
		}
	}
	void run_tasks()
	{
		for (;;)
		{
			const uint32_t i = next_.fetch_add(1);
			if (i >= num_tasks_)
				break;
			(*fn_)(i);
		}
	}
	void worker(int id, uint64_t seen)
	{
		for (;;)
		{
			{
				std::unique_lock<std::mutex> lk(mutex_);
				wake_.wait(lk, [&] { return generation_ != seen; });
				seen = generation_;
				if (stop_)
					return;
				if (id >= active_workers_)
					continue;
			}
			run_tasks();
			{
				std::lock_guard<std::mutex> lk(mutex_);
				if (--pending_ == 0)
					done_.notify_one();
			}
		}
	}

	std::mutex call_mutex_, mutex_;
	std::condition_variable wake_, done_;
	std::vector<std::thread> workers_;
	const std::function<void(uint32_t)>* fn_ = nullptr;
	uint32_t num_tasks_ = 0;
	std::atomic<uint32_t> next_{0};
	int active_workers_ = 0, pending_ = 0;
	uint64_t generation_ = 0;
	bool stop_ = false;
};

} // namespace oracle
