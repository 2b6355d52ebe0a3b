// progressbar.hpp — progress / cancel / live dose preview handle polled from another thread.
//
// Public surface of the reference's ProgressBar<T> (include/dxmc/progressbar.hpp:42-235). On the
// B200 path exposures complete in launch-sized groups: Transport calls exposureCompleted(n) after
// each launch and checks cancel() between launches. The preview image is a maximum-intensity
// projection of the dose buffer registered with setDoseData(); Transport refreshes that host
// buffer from the device accumulators between waves (at most four times a second) when a
// ProgressBar is attached.
#pragma once
#include "dxmc/floating.hpp"

#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <iomanip>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

namespace dxmc {
template <Floating T = double>
struct DoseProgressImageData {
    std::array<std::size_t, 2> dimensions = { 0, 0 };
    std::array<T, 2> spacing = { 0, 0 };
    std::vector<std::uint8_t> image;
};

template <Floating T>
class ProgressBar {
public:
    enum class Axis { X, Y, Z };

    ProgressBar() = default;
    ProgressBar(std::uint64_t totalExposures) { setTotalExposures(totalExposures); }

    void setTotalExposures(std::uint64_t totalExposures)
    {
        m_total = totalExposures;
        m_done = 0;
        m_start = std::chrono::system_clock::now();
    }
    void setPrefixMessage(const std::string& msg) { m_message = msg; }

    void exposureCompleted(std::uint64_t n = 1) // thread safe
    {
        m_done.fetch_add(n);
        const auto elapsed = std::chrono::duration_cast<std::chrono::seconds>(std::chrono::system_clock::now() - m_start).count();
        m_seconds.store(static_cast<T>(elapsed));
    }

    std::string getETA() const
    {
        const auto done = m_done.load();
        if (done == 0)
            return m_message + "ETA: estimating...";
        const auto total = m_total.load();
        const T remaining = m_seconds.load() / done * (total - done);
        const T percent = (T { 100 } * done) / total;
        std::stringstream ss;
        ss << std::fixed << std::setprecision(0) << m_message << "ETA: about ";
        if (remaining > 120)
            ss << remaining / 60 << " minutes";
        else
            ss << remaining << " seconds";
        ss << " [" << percent << "%]";
        return ss.str();
    }

    void setCancel(bool cancel) { m_cancel.store(cancel); }
    bool cancel() const { return m_cancel.load(); }

    void setPlaneNormal(Axis planeNormal) { m_axis = planeNormal; }
    void setDoseData(const T* doseData, const std::array<std::size_t, 3>& dimensions, const std::array<T, 3>& spacing)
    {
        std::scoped_lock guard(m_mutex);
        m_dose = doseData;
        m_dim = dimensions;
        m_spacing = spacing;
    }
    void clearDoseData()
    {
        std::scoped_lock guard(m_mutex);
        m_dose = nullptr;
        m_spacing.fill(T { 0 });
        m_dim.fill(0);
    }
    // the mutex also serialises Transport's refresh of the registered buffer
    std::mutex& doseMutex() { return m_mutex; }

    // 8-bit maximum-intensity projection along the chosen axis, scaled to the global maximum
    std::shared_ptr<DoseProgressImageData<T>> computeDoseProgressImage()
    {
        std::scoped_lock guard(m_mutex);
        if (!m_dose)
            return nullptr;
        const int normal = m_axis == Axis::X ? 0 : (m_axis == Axis::Y ? 1 : 2);
        const int u = normal == 0 ? 1 : 0;
        const int v = normal == 2 ? 1 : 2;
        auto img = std::make_shared<DoseProgressImageData<T>>();
        img->dimensions = { m_dim[u], m_dim[v] };
        img->spacing = { m_spacing[u], m_spacing[v] };
        std::vector<T> mip(m_dim[u] * m_dim[v], T { 0 });
        T globalMax = 0;
        std::array<std::size_t, 3> idx;
        for (idx[2] = 0; idx[2] < m_dim[2]; ++idx[2])
            for (idx[1] = 0; idx[1] < m_dim[1]; ++idx[1])
                for (idx[0] = 0; idx[0] < m_dim[0]; ++idx[0]) {
                    const T d = m_dose[idx[0] + m_dim[0] * (idx[1] + m_dim[1] * idx[2])];
                    T& m = mip[idx[u] + m_dim[u] * idx[v]];
                    m = std::max(m, d);
                    globalMax = std::max(globalMax, m);
                }
        const T scale = globalMax > 0 ? T { 255.0 } / globalMax : T { 0 }; // an all-zero buffer gives an all-zero image
        img->image.resize(mip.size());
        std::transform(mip.cbegin(), mip.cend(), img->image.begin(), [=](const T el) { return static_cast<std::uint8_t>(el * scale); });
        return img;
    }

private:
    std::atomic<std::uint64_t> m_total = 0;
    std::atomic<std::uint64_t> m_done = 0;
    std::chrono::system_clock::time_point m_start;
    std::atomic<T> m_seconds { 0 };
    std::string m_message;
    std::atomic<bool> m_cancel = false;
    std::mutex m_mutex;
    const T* m_dose = nullptr;
    std::array<std::size_t, 3> m_dim = { 0, 0, 0 };
    std::array<T, 3> m_spacing = { 1, 1, 1 };
    Axis m_axis = Axis::Y;
};
}
