/* dxmcb200_scene_monitor.hpp — shared implementation of dxs_transport_monitored.
 *
 * Included by BOTH implementations of the scene API (dxmclib_b200/host/scene_capi.cpp on the drop-in classes,
 * oracle/ref_harness.cpp on the unmodified reference headers): it only uses the public surface the two share,
 * Transport<T>::operator()(world, source, ProgressBar*, bool) and ProgressBar<T> (reference progressbar.hpp:42-235),
 * the way the reference's own callers do (validation.cpp:269-290: operator() on one thread, a second thread polling
 * getETA / computeDoseProgressImage / setCancel). */
#pragma once
#include "dxmcb200_scene.h"

#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>

namespace dxs_monitor {

inline double percentOf(const std::string& eta) /* "... [37%]" -> 37; no bracket (still estimating) -> 0 */
{
    const auto open = eta.rfind('[');
    const auto close = eta.rfind('%');
    if (open == std::string::npos || close == std::string::npos || close <= open)
        return 0.0;
    return std::atof(eta.substr(open + 1, close - open - 1).c_str());
}

template <typename ResultT, typename ProgressBarT, typename TransportT, typename WorldT, typename SourceT>
ResultT run(TransportT& transport, const WorldT& world, SourceT* source, bool useCalibration, double cancelAtPercent, dxs_progress_report* report)
{
    ProgressBarT bar;
    std::atomic<bool> finished { false };
    dxs_progress_report rep {};
    std::thread monitor([&]() {
        while (!finished.load()) {
            const std::string eta = bar.getETA();
            const double percent = percentOf(eta);
            if (percent > rep.percent_seen)
                rep.percent_seen = percent;
            if (cancelAtPercent > 0 && percent >= cancelAtPercent && !rep.cancelled) {
                bar.setCancel(true);
                rep.cancelled = 1;
            }
            if (!rep.cancelled) {
                if (const auto img = bar.computeDoseProgressImage()) {
                    ++rep.images_polled;
                    rep.image_width = static_cast<uint32_t>(img->dimensions[0]);
                    rep.image_height = static_cast<uint32_t>(img->dimensions[1]);
                    bool nonzero = false;
                    for (const auto px : img->image)
                        nonzero = nonzero || px != 0;
                    if (nonzero && img->image.size() == img->dimensions[0] * img->dimensions[1])
                        ++rep.images_nonzero;
                }
            }
            std::this_thread::sleep_for(std::chrono::milliseconds(2));
        }
    });
    ResultT result = transport(world, source, &bar, useCalibration);
    finished.store(true);
    monitor.join();
    const std::string eta = bar.getETA();
    rep.percent_final = percentOf(eta);
    std::strncpy(rep.eta, eta.c_str(), sizeof(rep.eta) - 1);
    if (report)
        *report = rep;
    return result;
}

} // namespace dxs_monitor
