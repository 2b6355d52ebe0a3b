// hostparallel.hpp — a plain thread-pool parallel-for for the host-side table builders.
//
// The reference marks its set-up loops std::execution::par_unseq (tube.hpp:191-208, attenuationinterpolator.hpp:51-59),
// which is serial unless libstdc++ finds TBB. The loops are element-wise independent, so running them on all host
// cores changes nothing in the values and takes the one-off source / table set-up from seconds to tenths of a second.
#pragma once
#include <algorithm>
#include <atomic>
#include <cstddef>
#include <thread>
#include <vector>

namespace dxmc {

namespace detail {
    // fn(i) for i in [0, n) on up to hardware_concurrency host threads; every i is independent
    template <typename F>
    inline void parallelFor(std::size_t n, F fn)
    {
        const std::size_t workers = std::min<std::size_t>(n, std::max(1u, std::thread::hardware_concurrency()));
        if (workers <= 1) {
            for (std::size_t i = 0; i < n; ++i)
                fn(i);
            return;
        }
        std::atomic<std::size_t> next { 0 };
        auto work = [&]() {
            for (std::size_t i = next.fetch_add(1); i < n; i = next.fetch_add(1))
                fn(i);
        };
        std::vector<std::thread> pool;
        for (std::size_t t = 1; t < workers; ++t)
            pool.emplace_back(work);
        work();
        for (auto& t : pool)
            t.join();
    }
}

}
