// transport.hpp — Transport<T>: run a Source through a World on a B200 and return per-voxel dose.
//
// Drop-in for the reference's Transport<T> (include/dxmc/transport.hpp:114-840): same setters,
// same call operator, same Result<T>, same output units and post-processing. What differs is the
// middle: where the reference spawns std::thread workers that pull exposures (parallellRun,
// :765-778), this class flattens the world, the look-up tables, the beam tables and ALL exposures
// into the POD structs of include/dxmcb200.h and hands them to the CUDA runtime in
// libdxmcb200.so. There is no CPU path: without a usable CUDA device the call THROWS std::runtime_error (the reference never
// throws from operator(); a silent all-zero dose from a machine without a GPU would be worse than an exception).
//
// Additions (no existing signature changed): setDevice(), setSeed(), setTracking(), lastStats().
// Precision: the device computes in single precision for every T. Transport<double> converts the world's densities and the
// look-up tables to float on upload and widens the results on download; sources and table builders run in T on the host.
#pragma once
#include "dxmc/attenuationlut.hpp"
#include "dxmc/exposure.hpp"
#include "dxmc/lowenergycorrectionmodel.hpp"
#include "dxmc/particle.hpp"
#include "dxmc/progressbar.hpp"
#include "dxmc/vectormath.hpp"
#include "dxmc/world.hpp"
#include "dxmcb200.h"

#include <algorithm>
#include <chrono>
#include <exception>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <optional>
#include <random>
#include <stdexcept>
#include <string>
#include <string_view>
#include <thread>
#include <vector>

namespace dxmc {

template <Floating T = double>
struct Result {
    std::vector<T> dose;
    std::vector<std::uint32_t> nEvents;
    std::vector<T> variance;
    std::uint64_t numberOfHistories { 0 };
    std::chrono::duration<float> simulationTime { 0 };
    std::string_view dose_units = "";

    Result() = default;
    Result(std::size_t size)
        : dose(size, T { 0 })
        , nEvents(size, 0)
        , variance(size, T { 0 })
    {
    }
    [[nodiscard]] std::vector<T> relativeError() const
    {
        std::vector<T> err(dose.size());
        for (std::size_t i = 0; i < dose.size(); ++i)
            err[i] = dose[i] > 0 ? std::sqrt(variance[i]) / dose[i] : 0;
        return err;
    }
};

template <Floating T>
class Source;

namespace detail {
    // device the next default-constructed Transport uses; the CT calibration run inherits the outer one
    inline int& currentDevice()
    {
        thread_local int device = 0;
        return device;
    }

    // Seed the calibration run of a CT source derives its own from: a Transport constructed while another one is collecting
    // (CTBaseSource::ctCalibration) takes a hash of the outer run's seed instead of drawing a fresh one, so that a seeded outer
    // run is reproducible end to end and the two runs never share per-history streams.
    inline std::optional<std::uint64_t>& inheritedSeed()
    {
        thread_local std::optional<std::uint64_t> seed;
        return seed;
    }
    inline std::uint64_t mixSeed(std::uint64_t z)
    {
        z += 0x9E3779B97F4A7C15ULL;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    }

    // NCCL communicators of a set of devices of this process, made once and kept (ncclCommInitAll costs a good fraction of a second)
    inline std::vector<void*> communicators(const std::vector<int>& devices)
    {
        static std::mutex guard;
        static std::map<std::vector<int>, std::vector<void*>> cache;
        std::scoped_lock lock(guard);
        auto it = cache.find(devices);
        if (it == cache.end()) {
            std::vector<void*> comms(devices.size(), nullptr);
            const int made = dxmcb200_comm_create(static_cast<int>(devices.size()), devices.data(), comms.data());
            if (made != DXMCB200_OK)
                throw std::runtime_error("dxmcb200: NCCL communicators for " + std::to_string(devices.size()) + " devices could not be made (status "
                    + std::to_string(made) + ")");
            it = cache.emplace(devices, std::move(comms)).first;
        }
        return it->second;
    }

    struct ContextDeleter {
        void operator()(dxmcb200_ctx* c) const { dxmcb200_destroy(c); }
    };
    using ContextPtr = std::unique_ptr<dxmcb200_ctx, ContextDeleter>;

    inline void check(dxmcb200_ctx* c, int status, const char* what)
    {
        if (status != DXMCB200_OK && status != DXMCB200_ERR_CANCELLED)
            throw std::runtime_error(std::string("dxmcb200: ") + what + " failed (" + std::to_string(status) + "): " + dxmcb200_last_error(c));
    }

    // everything the device needs besides the voxel arrays, in the layout of include/dxmcb200.h
    struct FlatTables {
        std::vector<float> knots, coefficients, maxCoefficients, rita, spline, shells;
        dxmcb200_luts luts {};

        struct Spectrum {
            std::vector<float> probs, energies;
            std::vector<std::uint32_t> alias;
        };
        struct Heel {
            dxmcb200_heel desc {};
            std::vector<float> weights;
        };
        struct Bowtie {
            std::vector<float> angles, weights;
        };
        std::vector<Spectrum> spectra;
        std::vector<Heel> heels;
        std::vector<Bowtie> bowties;
        std::vector<dxmcb200_exposure> exposures; // filled on the host only for sources that cannot describe themselves
        bool described = false; // the exposure table is generated on the device from `source`
        dxmcb200_source_params source {};
        std::vector<float> tubeCurrent; // source.aec_size slice intensities
        double maxWeight = 0; // largest possible photon birth weight
    };

    // DXMCB200_TRACE=1: phase timings of Transport::operator() on stderr
    struct PhaseTrace {
        bool on = false;
        std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
        PhaseTrace()
        {
            const char* env = std::getenv("DXMCB200_TRACE");
            on = env && env[0] == '1';
        }
        void operator()(const char* what)
        {
            if (!on)
                return;
            const auto now = std::chrono::steady_clock::now();
            std::fprintf(stderr, "[dxmcb200] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t).count());
            t = now;
        }
    };

    template <typename V>
    inline void appendFloats(std::vector<float>& out, const V& v)
    {
        for (const auto x : v)
            out.push_back(static_cast<float>(x));
    }
}

template <Floating T = double>
class Transport {
public:
    enum class OUTPUTMODE { EV_PER_HISTORY, DOSE };

    Transport()
        : m_nThreads(std::max<std::uint64_t>(std::thread::hardware_concurrency(), 1))
        , m_device(detail::currentDevice())
    {
        if (detail::inheritedSeed())
            m_seed = detail::mixSeed(*detail::inheritedSeed());
        else if (const char* env = std::getenv("DXMCB200_DEVICES")) { // "0,1,2,3" or "all"; not inherited by calibration runs
            std::vector<int> devices;
            if (std::string(env) == "all") {
                int n = 0;
                dxmcb200_device_count(&n);
                for (int k = 0; k < n; ++k)
                    devices.push_back(k);
            } else {
                for (const char* p = env; *p;) {
                    char* end = nullptr;
                    const long v = std::strtol(p, &end, 10);
                    if (end == p)
                        break;
                    devices.push_back(static_cast<int>(v));
                    p = *end ? end + 1 : end;
                }
            }
            setDevices(devices);
        }
    }

    // Copyable like the reference's Transport (validation.cpp passes it by value): a copy takes the settings and the
    // look-up tables, never the device context or the prepared run of the original.
    Transport(const Transport& other)
        : m_attenuationLut(other.m_attenuationLut)
        , m_nThreads(other.m_nThreads)
        , m_outputmode(other.m_outputmode)
        , m_lowenergyCorrection(other.m_lowenergyCorrection)
        , m_device(other.m_device)
        , m_seed(other.m_seed)
        , m_tracking(other.m_tracking)
        , m_brickMm(other.m_brickMm)
        , m_devices(other.m_devices)
    {
    }
    Transport& operator=(const Transport& other)
    {
        if (this != &other) {
            release();
            m_attenuationLut = other.m_attenuationLut;
            m_nThreads = other.m_nThreads;
            m_outputmode = other.m_outputmode;
            m_lowenergyCorrection = other.m_lowenergyCorrection;
            m_device = other.m_device;
            m_seed = other.m_seed;
            m_tracking = other.m_tracking;
            m_brickMm = other.m_brickMm;
            m_devices = other.m_devices;
            m_flat = detail::FlatTables {};
            m_totalExposures = m_histories = 0;
        }
        return *this;
    }
    Transport(Transport&&) = default;
    Transport& operator=(Transport&&) = default;

    // kept for source compatibility; the B200 path has no host workers
    void setNumberOfWorkers(std::uint64_t n) { m_nThreads = std::max(n, std::uint64_t { 1 }); }
    std::size_t numberOfWorkers() const { return m_nThreads; }
    void setLowEnergyCorrectionModel(LOWENERGYCORRECTION model) { m_lowenergyCorrection = model; }
    LOWENERGYCORRECTION lowEnergyCorrectionModel() const { return m_lowenergyCorrection; }
    void setOutputMode(OUTPUTMODE mode) { m_outputmode = mode; }
    OUTPUTMODE outputMode() const { return m_outputmode; }

    // ---- B200 additions
    void setDevice(int device)
    {
        m_device = device;
        m_devices.clear();
    }
    int device() const { return m_device; }
    // Several GPUs of this machine behind the one call (also DXMCB200_DEVICES=0,1,... or "all" in the environment): device k of n
    // transports exposures k, k + n, ... (world and tables replicated), the fixed-point grids are summed with ONE NCCL
    // reduce-scatter over voxel slices, and every device decodes and downloads its own slice into the Result. Integer sums:
    // the Result is bit-identical to the single-GPU one for the same seed.
    void setDevices(const std::vector<int>& devices)
    {
        m_devices = devices;
        if (!devices.empty())
            m_device = devices.front();
    }
    const std::vector<int>& devices() const { return m_devices; }
    // Master seed of the per-history counter streams. Unset (the default), every call draws a fresh seed from
    // std::random_device, like the reference's workers do (transport.hpp:749): repeated calls are independent samples and
    // can be averaged. setSeed makes the call reproducible bit for bit; seed() is the seed of the last (or next seeded) run.
    void setSeed(std::uint64_t seed) { m_seed = seed; }
    void clearSeed() { m_seed.reset(); }
    std::uint64_t seed() const { return m_seed ? *m_seed : m_runSeed; }
    // 0: the reference's Woodcock loop everywhere; 1 (default): + empty-space traversal through air bricks of about brickMm
    // (dxmcb200_set_tracking; -1 / 0 keep the library's defaults and environment overrides)
    void setTracking(int tracking, float brickMm = 0.0f)
    {
        m_tracking = tracking;
        m_brickMm = brickMm;
    }
    int tracking() const { return m_tracking; }
    const dxmcb200_stats& lastStats() const { return m_stats; }

    template <typename U>
        requires std::is_base_of_v<World<T>, U>
    Result<T> operator()(const U& world, Source<T>* source, ProgressBar<T>* progressbar = nullptr, bool useSourceDoseCalibration = true)
    {
        detail::PhaseTrace trace;
        if (m_devices.size() > 1)
            return runOnDevices(world, source, progressbar, useSourceDoseCalibration);
        if (!prepare(world, source))
            return Result<T>(world.size());
        trace("prepare (total)");
        Result<T> result(0);
        bool completed = false;
        if (progressbar) { // the progress image reads result.dose while the run is in flight (progressbar.hpp:86-109)
            result = Result<T>(world.size());
            trace("Result allocation");
            completed = run(0, m_totalExposures, progressbar, &result, &world);
        } else { // otherwise the three result arrays (zero-filled, page-faulted) are made while the GPU works
            std::thread allocation([&]() { result = Result<T>(world.size()); });
            struct Join {
                std::thread& t;
                ~Join() { t.join(); }
            } join { allocation };
            completed = run<U>(0, m_totalExposures, nullptr, nullptr, nullptr);
        }
        result.numberOfHistories = m_histories;
        result.simulationTime = m_lastRunTime;
        trace("run");
        if (completed)
            collect(world, source, result, useSourceDoseCalibration, progressbar);
        else { // cancelled: all zeros, like a cancelled reference run (transport.hpp:176-185)
            std::fill(result.dose.begin(), result.dose.end(), T { 0 }); // may hold a live preview of the partial dose
            result.numberOfHistories = 0;
        }
        trace("collect");
        release();
        trace("release");
        return result;
    }

    // ---- the three phases of operator(), public so that a caller can keep the world, tables and
    // exposures resident on the GPU across runs (bench.py, multi-GPU drivers) -------------------------

    // validate, build the look-up tables on the host, flatten and upload everything. Returns false for the
    // inputs the reference answers with an all-zero Result (invalid world / null or invalid source).
    // historiesAllRanks: total over all GPUs of a sharded run (sizes the fixed-point scale); 0 = this source.
    template <typename U>
        requires std::is_base_of_v<World<T>, U>
    bool prepare(const U& world, Source<T>* source, std::uint64_t historiesAllRanks = 0)
    {
        release();
        detail::PhaseTrace trace;
        if (!world.isValid() || !source)
            return false;
        source->updateFromWorld(world);
        source->validate();
        if (!source->isValid())
            return false;
        trace("  source update/validate");
        m_histories = source->historiesPerExposure() * source->totalExposures();
        m_totalExposures = source->totalExposures();

        dxmcb200_ctx* raw = nullptr;
        const int created = dxmcb200_create(m_device, &raw);
        if (created != DXMCB200_OK)
            throw std::runtime_error("dxmcb200: no usable CUDA device " + std::to_string(m_device) + " (status " + std::to_string(created)
                + "); this library has no CPU fallback");
        m_ctx.reset(raw);
        if (m_tracking >= 0)
            detail::check(m_ctx.get(), dxmcb200_set_tracking(m_ctx.get(), m_tracking, m_brickMm), "set_tracking");
        if (m_seed) {
            m_runSeed = *m_seed;
        } else {
            std::random_device entropy;
            m_runSeed = (static_cast<std::uint64_t>(entropy()) << 32) ^ entropy();
        }
        trace("  create context");
        // The voxel grid goes to the device first, on a helper thread, while this thread builds the per-material
        // tables (form-factor samplers, scatter functions, shells: independent of the grid). The one thing the
        // Woodcock majorant needs from the grid, the per-material maximum density, then comes back from the device
        // instead of a host pass over all voxels (reference attenuationinterpolator.hpp:48-59).
        std::vector<T> maxDensity;
        std::exception_ptr uploadError;
        std::thread upload([&]() {
            try {
                uploadWorld(m_ctx.get(), world);
                std::vector<float> md(std::min<std::size_t>(world.materialMap().size(), 256), 0.0f);
                detail::check(m_ctx.get(), dxmcb200_material_max_density(m_ctx.get(), static_cast<std::uint32_t>(md.size()), md.data()),
                    "material_max_density");
                maxDensity.assign(md.begin(), md.end());
            } catch (...) {
                uploadError = std::current_exception();
            }
        });
        try {
            m_attenuationLut.generate(world.materialMap(), source->maxPhotonEnergyProduced(), T { 1 }, false);
        } catch (...) {
            upload.join();
            throw;
        }
        upload.join();
        if (uploadError)
            std::rethrow_exception(uploadError);
        trace("  upload world || material tables");
        m_attenuationLut.generateAttenuation(world, maxDensity);
        trace("  attenuation fits + majorant");

        m_flat = detail::FlatTables {};
        flattenLuts(m_flat);
        if (!describeSource(world, *source, m_flat))
            flattenExposures(world, *source, m_totalExposures, m_flat);
        trace("  flatten tables/exposures");
        m_energyBits = 20;
        m_energySqBits = 10;
        dxmcb200_suggest_fixed_point(std::max(historiesAllRanks, m_histories), m_flat.maxWeight * static_cast<double>(source->maxPhotonEnergyProduced()),
            &m_energyBits, &m_energySqBits);
        uploadRun(m_ctx.get());
        return true;
    }

    // transport exposures [begin, end) into the device accumulators; false when cancelled
    template <typename U = World<T>>
    bool run(std::uint64_t begin, std::uint64_t end, ProgressBar<T>* progressbar = nullptr, Result<T>* result = nullptr, const U* world = nullptr)
    {
        if (!m_ctx)
            throw std::runtime_error("dxmcb200: Transport::run called before prepare");
        if (progressbar) {
            progressbar->setTotalExposures(end - begin);
            if (result && world)
                progressbar->setDoseData(result->dose.data(), world->dimensions(), world->spacing());
        }
        struct Progress {
            ProgressBar<T>* bar;
            dxmcb200_ctx* ctx;
            T* liveDose; // the buffer registered with ProgressBar::setDoseData, or null
            std::size_t voxels;
            std::uint64_t reported = 0;
            volatile int cancel = 0;
            std::chrono::steady_clock::time_point lastRefresh {};
            std::vector<float> staging; // T = double only
        } progress { progressbar, m_ctx.get(), (progressbar && result) ? result->dose.data() : nullptr, result ? result->dose.size() : 0 };
        // Called by the runtime on this thread between waves. The reference's workers add into result.dose directly, so a
        // GUI thread polling ProgressBar::computeDoseProgressImage sees the dose build up (progressbar.hpp:86-109); here the
        // registered buffer is refreshed from the device accumulators (raw keV sums, like the reference's live buffer) at
        // most four times a second: one decode pass + download, about 1 % of the run at 512x512x400.
        auto callback = [](std::uint64_t done, void* user) {
            auto* p = static_cast<Progress*>(user);
            if (!p->bar)
                return;
            p->bar->exposureCompleted(done - p->reported);
            p->reported = done;
            p->cancel = p->bar->cancel() ? 1 : 0;
            const auto now = std::chrono::steady_clock::now();
            if (p->liveDose && !p->cancel && now - p->lastRefresh >= std::chrono::milliseconds(250)) {
                std::scoped_lock guard(p->bar->doseMutex());
                if constexpr (std::is_same_v<T, float>) {
                    dxmcb200_get_result(p->ctx, 2, 1, 1.0f, p->liveDose, nullptr, nullptr);
                } else {
                    p->staging.resize(p->voxels);
                    if (dxmcb200_get_result(p->ctx, 2, 1, 1.0f, p->staging.data(), nullptr, nullptr) == DXMCB200_OK)
                        std::copy(p->staging.begin(), p->staging.end(), p->liveDose);
                }
                p->lastRefresh = std::chrono::steady_clock::now();
            }
        };
        if (progressbar && progressbar->cancel())
            progress.cancel = 1;
        const auto start = std::chrono::system_clock::now();
        const int ran = dxmcb200_run_range(m_ctx.get(), begin, end, static_cast<int>(m_lowenergyCorrection), m_runSeed, &progress.cancel, callback, &progress);
        m_lastRunTime = std::chrono::system_clock::now() - start;
        if (result)
            result->simulationTime = m_lastRunTime;
        detail::check(m_ctx.get(), ran, "run");
        dxmcb200_get_stats(m_ctx.get(), &m_stats);
        if (progressbar) {
            progressbar->clearDoseData();
            if (progressbar->cancel())
                return false;
        }
        return ran != DXMCB200_ERR_CANCELLED;
    }

    // multi-GPU interleaving: transport exposures first, first + stride, ... (count of them); see dxmcb200_run_strided
    void runStrided(std::uint64_t first, std::uint64_t stride, std::uint64_t count)
    {
        if (!m_ctx)
            throw std::runtime_error("dxmcb200: Transport::runStrided called before prepare");
        const auto start = std::chrono::system_clock::now();
        const int ran = dxmcb200_run_strided(m_ctx.get(), first, stride, count, static_cast<int>(m_lowenergyCorrection), m_runSeed);
        m_lastRunTime = std::chrono::system_clock::now() - start;
        detail::check(m_ctx.get(), ran, "run_strided");
        dxmcb200_get_stats(m_ctx.get(), &m_stats);
    }

    // decode the accumulators with the reference's post-processing for the current output mode
    template <typename U>
        requires std::is_base_of_v<World<T>, U>
    void collect(const U& world, Source<T>* source, Result<T>& result, bool useSourceDoseCalibration = true, ProgressBar<T>* progressbar = nullptr)
    {
        if (!m_ctx)
            throw std::runtime_error("dxmcb200: Transport::collect called before prepare");
        if (result.dose.size() != world.size())
            result = Result<T>(world.size());
        if (result.numberOfHistories == 0)
            result.numberOfHistories = m_histories;
        int mode = 0;
        float calibration = 1.0f;
        if (m_outputmode == OUTPUTMODE::DOSE) {
            mode = 1;
            if (useSourceDoseCalibration) {
                calibration = static_cast<float>(calibrationValue(source, progressbar));
                result.dose_units = "mGy";
            } else {
                result.dose_units = "keV/kg";
            }
        } else {
            result.dose_units = "eV/history";
        }
        download(m_ctx.get(), mode, result, calibration);
    }

    // collect() straight into caller-owned arrays (any may be null), without the intermediate Result: what a multi-GPU
    // driver calls on the rank that keeps the summed grids. Returns the dose units.
    template <typename U>
        requires std::is_base_of_v<World<T>, U> && std::is_same_v<T, float>
    std::string_view collectInto(const U& world, Source<T>* source, std::uint64_t histories, float* dose, std::uint32_t* nEvents, float* variance,
        bool useSourceDoseCalibration = true)
    {
        if (!m_ctx)
            throw std::runtime_error("dxmcb200: Transport::collectInto called before prepare");
        (void)world;
        int mode = 0;
        float calibration = 1.0f;
        std::string_view units = "eV/history";
        if (m_outputmode == OUTPUTMODE::DOSE) {
            mode = 1;
            units = "keV/kg";
            if (useSourceDoseCalibration) {
                calibration = static_cast<float>(calibrationValue(source, nullptr));
                units = "mGy";
            }
        }
        detail::check(m_ctx.get(), dxmcb200_get_result(m_ctx.get(), mode, histories ? histories : m_histories, calibration, dose, nEvents, variance),
            "get_result");
        return units;
    }

    // device context of a prepared Transport, for direct C-ABI calls (accumulator address, stats, raw grids)
    dxmcb200_ctx* context() const { return m_ctx.get(); }
    std::uint64_t preparedExposures() const { return m_totalExposures; }
    std::uint64_t preparedHistories() const { return m_histories; }
    std::chrono::duration<double> lastRunTime() const { return m_lastRunTime; } // what Result::simulationTime reports
    void release() { m_ctx.reset(); }

    const AttenuationLut<T>& attenuationLut() const { return m_attenuationLut; }
    AttenuationLut<T>& attenuationLut() { return m_attenuationLut; }

protected:
    // everything a device needs besides the voxel grid: look-up tables, beam tables, fixed-point scale, the exposure table
    void uploadRun(dxmcb200_ctx* ctx) const
    {
        detail::check(ctx, dxmcb200_set_luts(ctx, &m_flat.luts), "set_luts");
        uploadBeamTables(ctx, m_flat);
        detail::check(ctx, dxmcb200_set_fixed_point(ctx, m_energyBits, m_energySqBits), "set_fixed_point");
        if (m_flat.described) // all exposures evaluated on the device from the source's parameter block
            detail::check(ctx, dxmcb200_generate_exposures(ctx, &m_flat.source, m_flat.tubeCurrent.data(), nullptr), "generate_exposures");
        else if (!m_flat.exposures.empty())
            detail::check(ctx, dxmcb200_upload_exposures(ctx, m_flat.exposures.data(), m_flat.exposures.size()), "upload_exposures");
    }

    // operator() over several GPUs of this process, one host thread per device
    template <typename U>
    Result<T> runOnDevices(const U& world, Source<T>* source, ProgressBar<T>* progressbar, bool useSourceDoseCalibration)
    {
        const int n = static_cast<int>(m_devices.size());
        m_device = m_devices.front();
        // the other devices take their copy of the voxel grid while this thread validates, builds the tables and prepares device 0
        std::vector<detail::ContextPtr> peers(static_cast<std::size_t>(n - 1));
        std::vector<std::exception_ptr> failure(static_cast<std::size_t>(n));
        std::vector<std::thread> uploads;
        const bool worldUsable = world.isValid() && source;
        for (int k = 1; worldUsable && k < n; ++k)
            uploads.emplace_back([&, k]() {
                try {
                    dxmcb200_ctx* raw = nullptr;
                    const int created = dxmcb200_create(m_devices[static_cast<std::size_t>(k)], &raw);
                    if (created != DXMCB200_OK)
                        throw std::runtime_error("dxmcb200: no usable CUDA device " + std::to_string(m_devices[static_cast<std::size_t>(k)]));
                    peers[static_cast<std::size_t>(k - 1)].reset(raw);
                    if (m_tracking >= 0)
                        detail::check(raw, dxmcb200_set_tracking(raw, m_tracking, m_brickMm), "set_tracking");
                    uploadWorld(raw, world);
                } catch (...) {
                    failure[static_cast<std::size_t>(k)] = std::current_exception();
                }
            });
        bool prepared = false;
        try {
            prepared = prepare(world, source);
        } catch (...) {
            failure[0] = std::current_exception();
        }
        for (auto& t : uploads)
            t.join();
        for (const auto& f : failure)
            if (f)
                std::rethrow_exception(f);
        if (!prepared)
            return Result<T>(world.size());
        const auto comms = detail::communicators(m_devices);
        auto contextOf = [&](int k) { return k == 0 ? m_ctx.get() : peers[static_cast<std::size_t>(k - 1)].get(); };

        // the three result arrays (zero-filled, page-faulted: 0.3 s at 512^3) are made while the GPUs work
        Result<T> result(0);
        std::vector<float> dose32, variance32; // T = double: the devices deliver float
        std::thread allocation([&]() {
            result = Result<T>(world.size());
            if constexpr (!std::is_same_v<T, float>) {
                dose32.resize(world.size());
                variance32.resize(world.size());
            }
        });
        struct Join {
            std::thread& t;
            ~Join()
            {
                if (t.joinable())
                    t.join();
            }
        } joinAllocation { allocation };
        int mode = 0;
        float calibration = 1.0f;
        std::string_view units = "eV/history";
        if (m_outputmode == OUTPUTMODE::DOSE) {
            mode = 1;
            units = "keV/kg";
            if (useSourceDoseCalibration) { // before the run: a CT source calibrates with a Transport of its own on device 0
                calibration = static_cast<float>(calibrationValue(source, progressbar));
                units = "mGy";
            }
        }
        if (progressbar)
            progressbar->setTotalExposures(m_totalExposures);
        struct Shared {
            ProgressBar<T>* bar;
            volatile int cancel = 0;
            std::uint64_t reported = 0;
            int ranks = 1;
        } shared { progressbar };
        shared.ranks = n;
        if (progressbar && progressbar->cancel())
            shared.cancel = 1;
        auto callback = [](std::uint64_t done, void* user) { // rank 0 only: its share of the exposures stands for all ranks
            auto* p = static_cast<Shared*>(user);
            if (!p->bar)
                return;
            p->bar->exposureCompleted((done - p->reported) * static_cast<std::uint64_t>(p->ranks));
            p->reported = done;
            if (p->bar->cancel())
                p->cancel = 1;
        };
        std::vector<int> status(static_cast<std::size_t>(n), DXMCB200_OK);
        const auto start = std::chrono::system_clock::now();
        auto rank = [&](int k) {
            try {
                dxmcb200_ctx* ctx = contextOf(k);
                if (k > 0)
                    uploadRun(ctx);
                const std::uint64_t count = static_cast<std::uint64_t>(k) < m_totalExposures ? (m_totalExposures - static_cast<std::uint64_t>(k) + n - 1) / n : 0;
                int ran = dxmcb200_run_strided_monitored(ctx, static_cast<std::uint64_t>(k), static_cast<std::uint64_t>(n), count,
                    static_cast<int>(m_lowenergyCorrection), m_runSeed, &shared.cancel, k == 0 ? +callback : nullptr, &shared);
                status[static_cast<std::size_t>(k)] = ran;
                detail::check(ctx, ran, "run_strided");
            } catch (...) {
                failure[static_cast<std::size_t>(k)] = std::current_exception();
            }
        };
        auto onAllDevices = [&](auto&& work) {
            std::vector<std::thread> threads;
            for (int k = 1; k < n; ++k)
                threads.emplace_back(work, k);
            work(0);
            for (auto& t : threads)
                t.join();
            for (const auto& f : failure)
                if (f)
                    std::rethrow_exception(f);
        };
        onAllDevices(rank);
        m_lastRunTime = std::chrono::system_clock::now() - start;
        allocation.join();
        result.numberOfHistories = m_histories;
        result.dose_units = units;
        result.simulationTime = m_lastRunTime;
        float* dose = nullptr;
        float* variance = nullptr;
        if constexpr (std::is_same_v<T, float>) {
            dose = result.dose.data();
            variance = result.variance.data();
        } else {
            dose = dose32.data();
            variance = variance32.data();
        }
        dxmcb200_get_stats(m_ctx.get(), &m_stats);
        const bool cancelled = shared.cancel != 0 || std::any_of(status.begin(), status.end(), [](int s) { return s == DXMCB200_ERR_CANCELLED; });
        if (cancelled) {
            result.numberOfHistories = 0;
        } else {
            onAllDevices([&](int k) {
                try {
                    dxmcb200_ctx* ctx = contextOf(k);
                    detail::check(ctx,
                        dxmcb200_reduce_collect(ctx, comms[static_cast<std::size_t>(k)], k, n, mode, m_histories, calibration, dose, result.nEvents.data(), variance),
                        "reduce_collect");
                } catch (...) {
                    failure[static_cast<std::size_t>(k)] = std::current_exception();
                }
            });
            if constexpr (!std::is_same_v<T, float>) {
                std::copy(dose32.begin(), dose32.end(), result.dose.begin());
                std::copy(variance32.begin(), variance32.end(), result.variance.begin());
            }
        }
        peers.clear();
        release();
        return result;
    }

    // Source::getCalibrationValue with this run's device and a seed derived from this run's for any Transport it constructs
    T calibrationValue(Source<T>* source, ProgressBar<T>* progressbar) const
    {
        struct Scope {
            int device = detail::currentDevice();
            std::optional<std::uint64_t> seed = detail::inheritedSeed();
            ~Scope()
            {
                detail::currentDevice() = device;
                detail::inheritedSeed() = seed;
            }
        } restore;
        detail::currentDevice() = m_device;
        detail::inheritedSeed() = m_runSeed;
        return source->getCalibrationValue(m_lowenergyCorrection, progressbar);
    }

    // The source as a parameter block for device-side exposure generation, with the beam tables its tubes point to collected
    // once each. False when the source cannot describe itself (a user-defined Source that overrides getExposure only).
    template <typename U>
    bool describeSource(const U& world, const Source<T>& source, detail::FlatTables& f) const
    {
        model::SourceParams<T> block;
        const T* profile = nullptr;
        if (!source.describe(block, profile))
            return false;
        model::SourceParams<float> p;
        p.motion = block.motion;
        p.tubes = block.tubes;
        p.exposures = block.exposures;
        p.histories = block.histories;
        const auto narrow = [](const auto& from, auto& to) {
            for (std::size_t k = 0; k < sizeof(to) / sizeof(to[0]); ++k)
                to[k] = static_cast<float>(from[k]);
        };
        narrow(block.position, p.position);
        narrow(block.cosines, p.cosines);
        narrow(block.collimation, p.collimation);
        narrow(block.sdd, p.sdd);
        narrow(block.fov, p.fov);
        narrow(block.startAngle, p.startAngle);
        narrow(block.tubeWeight, p.tubeWeight);
        p.monoEnergy = static_cast<float>(block.monoEnergy);
        p.focalOffset = static_cast<float>(block.focalOffset);
        p.orbitFullTurn = block.orbitFullTurn;
        p.orbitStep = static_cast<float>(block.orbitStep);
        p.beamWidth = static_cast<float>(block.beamWidth);
        p.angleStep = static_cast<float>(block.angleStep);
        p.pitch = static_cast<float>(block.pitch);
        p.tableStep = static_cast<float>(block.tableStep);
        p.tilt = static_cast<float>(block.tilt);
        p.scanLength = static_cast<float>(block.scanLength);
        p.xcare = block.xcare;
        p.xcareAngle = static_cast<float>(block.xcareAngle);
        p.xcareSpan = static_cast<float>(block.xcareSpan);
        p.xcareRamp = static_cast<float>(block.xcareRamp);
        p.xcareLow = static_cast<float>(block.xcareLow);
        p.aecSize = block.aecSize;
        p.aecMin = static_cast<float>(block.aecMin);
        p.aecMax = static_cast<float>(block.aecMax);
        p.aecStep = static_cast<float>(block.aecStep);
        p.align = 1;
        for (std::size_t k = 0; k < 6; ++k)
            p.worldCosines[k] = static_cast<float>(world.directionCosines()[k]);
        double currentMax = 1.0; // largest product of the two tube-current modulations
        if (block.aecSize) {
            f.tubeCurrent.assign(profile, profile + block.aecSize);
            currentMax = std::max(0.0, static_cast<double>(*std::max_element(f.tubeCurrent.begin(), f.tubeCurrent.end())));
        }
        if (block.xcare) {
            constexpr double twoPi = 6.283185307179586;
            currentMax *= std::max(1.0, (twoPi - block.xcareSpan * block.xcareLow + block.xcareLow * block.xcareRamp) / (twoPi - block.xcareSpan + block.xcareRamp));
        }
        for (std::uint32_t t = 0; t < block.tubes; ++t) {
            const auto tables = source.beamTables(t);
            double w = block.tubes == 2 ? std::abs(static_cast<double>(block.tubeWeight[t])) : 1.0;
            p.spectrum[t] = p.heel[t] = p.bowtie[t] = -1;
            if (tables.specter) {
                p.spectrum[t] = static_cast<std::int32_t>(f.spectra.size());
                f.spectra.push_back(flatSpectrum(*tables.specter));
            }
            if (tables.heel) {
                p.heel[t] = static_cast<std::int32_t>(f.heels.size());
                f.heels.push_back(flatHeel(*tables.heel));
                w *= std::max(1.0, static_cast<double>(*std::max_element(f.heels.back().weights.begin(), f.heels.back().weights.end())));
            }
            if (tables.fan) {
                p.bowtie[t] = static_cast<std::int32_t>(f.bowties.size());
                f.bowties.push_back(flatFan(*tables.fan));
                w *= std::max(1.0, static_cast<double>(*std::max_element(f.bowties.back().weights.begin(), f.bowties.back().weights.end())));
            }
            f.maxWeight = std::max(f.maxWeight, w * currentMax);
        }
        if (f.maxWeight <= 0)
            f.maxWeight = 1;
        static_assert(sizeof(p) == sizeof(f.source));
        std::memcpy(&f.source, &p, sizeof(p));
        f.described = true;
        return true;
    }

    static detail::FlatTables::Spectrum flatSpectrum(const SpecterDistribution<T>& s)
    {
        detail::FlatTables::Spectrum fs;
        detail::appendFloats(fs.probs, s.probabilityData());
        detail::appendFloats(fs.energies, s.energies());
        for (auto a : s.aliasingData())
            fs.alias.push_back(static_cast<std::uint32_t>(a));
        return fs;
    }
    static detail::FlatTables::Heel flatHeel(const HeelFilter<T>& h)
    {
        detail::FlatTables::Heel fh;
        fh.desc.energy_start = static_cast<float>(h.energyStart());
        fh.desc.energy_step = static_cast<float>(h.energyStep());
        fh.desc.energy_size = static_cast<std::uint32_t>(h.energySize());
        fh.desc.angle_start = static_cast<float>(h.angleStart());
        fh.desc.angle_step = static_cast<float>(h.angleStep());
        fh.desc.angle_size = static_cast<std::uint32_t>(h.angleSize());
        detail::appendFloats(fh.weights, h.weights());
        return fh;
    }
    static detail::FlatTables::Bowtie flatFan(const BeamFilter<T>& b)
    {
        detail::FlatTables::Bowtie fb;
        if (const auto* bt = dynamic_cast<const BowTieFilter<T>*>(&b)) {
            for (const auto& [angle, weight] : bt->data()) {
                fb.angles.push_back(static_cast<float>(angle));
                fb.weights.push_back(static_cast<float>(weight));
            }
        } else { // any other BeamFilter is tabulated on |angle| in [0, pi/2] (symmetric filters only)
            constexpr int n = 2048;
            for (int k = 0; k < n; ++k) {
                const T a = (PI_VAL<T>() / 2) * k / (n - 1);
                fb.angles.push_back(static_cast<float>(a));
                fb.weights.push_back(static_cast<float>(b.sampleIntensityWeight(a)));
            }
        }
        return fb;
    }

    void flattenLuts(detail::FlatTables& f) const
    {
        const auto& lut = m_attenuationLut;
        const auto& ip = lut.attenuationData();
        const std::size_t nMat = lut.formFactorSamplers().size();
        detail::appendFloats(f.knots, ip.knots());
        detail::appendFloats(f.coefficients, ip.coefficients());
        detail::appendFloats(f.maxCoefficients, ip.maxCoefficients());
        for (std::size_t m = 0; m < nMat; ++m) {
            const auto& r = lut.formFactorSamplers()[m];
            detail::appendFloats(f.rita, r.x());
            detail::appendFloats(f.rita, r.e());
            detail::appendFloats(f.rita, r.a());
            detail::appendFloats(f.rita, r.b());
            const auto& s = lut.scatterFunctions()[m];
            detail::appendFloats(f.spline, s.coefficients());
            f.spline.push_back(static_cast<float>(s.start()));
            f.spline.push_back(static_cast<float>(s.step()));
            f.spline.push_back(static_cast<float>(s.stop()));
            for (const auto& sh : lut.electronShellConfiguration(m)) {
                const T row[DXMCB200_SHELL_FLOATS] = { sh.bindingEnergy, sh.numberElectrons, sh.hartreeFockOrbital_0, sh.photoIonizationProbability,
                    sh.fluorescenceYield, sh.fluorLineProbabilities[0], sh.fluorLineProbabilities[1], sh.fluorLineProbabilities[2], sh.fluorLineEnergies[0],
                    sh.fluorLineEnergies[1], sh.fluorLineEnergies[2] };
                detail::appendFloats(f.shells, row);
            }
        }
        f.luts.n_materials = static_cast<std::uint32_t>(nMat);
        f.luts.n_segments = static_cast<std::uint32_t>(ip.resolution());
        f.luts.linear_index = static_cast<std::uint32_t>(ip.linearIndex());
        f.luts.linear_step = static_cast<float>(ip.linearStep());
        f.luts.linear_energy = static_cast<float>(ip.linearEnergy());
        f.luts.knots = f.knots.data();
        f.luts.coefficients = f.coefficients.data();
        f.luts.max_coefficients = f.maxCoefficients.data();
        f.luts.rita = f.rita.data();
        f.luts.spline = f.spline.data();
        f.luts.shells = f.shells.data();
    }

    // getExposure(i) + alignToDirectionCosines for every exposure (reference transport.hpp:756-757),
    // with the beam tables the exposures point to collected once each
    template <typename U>
    void flattenExposures(const U& world, const Source<T>& source, std::uint64_t totalExposures, detail::FlatTables& f) const
    {
        std::map<const void*, std::int32_t> spectrumIdx, heelIdx, bowtieIdx;
        std::vector<double> heelMax, bowtieMax; // per-table largest weight factor
        f.exposures.reserve(totalExposures);
        for (std::uint64_t i = 0; i < totalExposures; ++i) {
            auto e = source.getExposure(i);
            e.alignToDirectionCosines(world.directionCosines());
            dxmcb200_exposure pod {};
            for (int k = 0; k < 3; ++k) {
                pod.position[k] = static_cast<float>(e.position()[k]);
                pod.beam_direction[k] = static_cast<float>(e.beamDirection()[k]);
            }
            for (int k = 0; k < 6; ++k)
                pod.cosines[k] = static_cast<float>(e.directionCosines()[k]);
            for (int k = 0; k < 4; ++k)
                pod.collimation[k] = static_cast<float>(e.collimationAngles()[k]);
            pod.weight = static_cast<float>(e.beamIntensityWeight());
            pod.mono_energy = static_cast<float>(e.monoenergeticPhotonEnergy());
            pod.histories = e.numberOfHistories();
            pod.spectrum = pod.heel = pod.bowtie = -1;
            double w = std::abs(static_cast<double>(pod.weight));

            if (const auto* s = e.specterDistribution()) {
                auto [it, isNew] = spectrumIdx.try_emplace(s, static_cast<std::int32_t>(f.spectra.size()));
                if (isNew)
                    f.spectra.push_back(flatSpectrum(*s));
                pod.spectrum = it->second;
            }
            if (const auto* h = e.heelFilter()) {
                auto [it, isNew] = heelIdx.try_emplace(h, static_cast<std::int32_t>(f.heels.size()));
                if (isNew) {
                    f.heels.push_back(flatHeel(*h));
                    heelMax.push_back(*std::max_element(f.heels.back().weights.begin(), f.heels.back().weights.end()));
                }
                pod.heel = it->second;
                w *= std::max(1.0, heelMax[it->second]);
            }
            if (const auto* b = e.beamFilter()) {
                auto [it, isNew] = bowtieIdx.try_emplace(b, static_cast<std::int32_t>(f.bowties.size()));
                if (isNew) {
                    f.bowties.push_back(flatFan(*b));
                    bowtieMax.push_back(*std::max_element(f.bowties.back().weights.begin(), f.bowties.back().weights.end()));
                }
                pod.bowtie = it->second;
                w *= std::max(1.0, bowtieMax[it->second]);
            }
            f.maxWeight = std::max(f.maxWeight, w);
            f.exposures.push_back(pod);
        }
        if (f.maxWeight <= 0)
            f.maxWeight = 1;
    }

    template <typename U>
    void uploadWorld(dxmcb200_ctx* ctx, const U& world) const
    {
        dxmcb200_world w {};
        for (int i = 0; i < 3; ++i) {
            w.dim[i] = world.dimensions()[i];
            w.spacing[i] = static_cast<float>(world.spacing()[i]);
        }
        for (int i = 0; i < 6; ++i)
            w.extent_safe[i] = static_cast<float>(world.matrixExtentSafe()[i]);
        std::vector<float> converted;
        if constexpr (std::is_same_v<T, float>) {
            w.density = world.densityArray()->data();
        } else {
            converted.assign(world.densityArray()->begin(), world.densityArray()->end());
            w.density = converted.data();
        }
        w.material = world.materialIndexArray()->data();
        w.measurement = world.measurementMapArray() ? world.measurementMapArray()->data() : nullptr;
        detail::check(ctx, dxmcb200_set_world(ctx, &w), "set_world");
    }

    void uploadBeamTables(dxmcb200_ctx* ctx, const detail::FlatTables& f) const
    {
        std::vector<dxmcb200_spectrum> s;
        std::vector<dxmcb200_heel> h;
        std::vector<dxmcb200_bowtie> b;
        for (const auto& x : f.spectra)
            s.push_back({ static_cast<std::uint32_t>(x.probs.size()), x.probs.data(), x.alias.data(), x.energies.data() });
        for (const auto& x : f.heels) {
            auto d = x.desc;
            d.weights = x.weights.data();
            h.push_back(d);
        }
        for (const auto& x : f.bowties)
            b.push_back({ static_cast<std::uint32_t>(x.angles.size()), x.angles.data(), x.weights.data() });
        detail::check(ctx,
            dxmcb200_set_beam_tables(ctx, static_cast<std::uint32_t>(s.size()), s.data(), static_cast<std::uint32_t>(h.size()), h.data(),
                static_cast<std::uint32_t>(b.size()), b.data()),
            "set_beam_tables");
    }

    void download(dxmcb200_ctx* ctx, int mode, Result<T>& result, float calibration) const
    {
        if constexpr (std::is_same_v<T, float>) {
            detail::check(ctx,
                dxmcb200_get_result(ctx, mode, result.numberOfHistories, calibration, result.dose.data(), result.nEvents.data(), result.variance.data()),
                "get_result");
        } else {
            std::vector<float> dose(result.dose.size()), variance(result.variance.size());
            detail::check(ctx, dxmcb200_get_result(ctx, mode, result.numberOfHistories, calibration, dose.data(), result.nEvents.data(), variance.data()),
                "get_result");
            std::copy(dose.begin(), dose.end(), result.dose.begin());
            std::copy(variance.begin(), variance.end(), result.variance.begin());
        }
    }

private:
    AttenuationLut<T> m_attenuationLut;
    std::uint64_t m_nThreads;
    OUTPUTMODE m_outputmode = OUTPUTMODE::DOSE;
    LOWENERGYCORRECTION m_lowenergyCorrection = LOWENERGYCORRECTION::LIVERMORE;
    int m_device = 0;
    std::optional<std::uint64_t> m_seed; // unset: a fresh seed per call
    std::uint64_t m_runSeed = 0; // seed of the prepared / last run
    int m_tracking = -1; // -1: library default
    float m_brickMm = 0.0f;
    std::vector<int> m_devices; // more than one entry: operator() runs on all of them
    int m_energyBits = 20, m_energySqBits = 10;
    dxmcb200_stats m_stats {};
    std::chrono::duration<float> m_lastRunTime {};
    detail::ContextPtr m_ctx;
    detail::FlatTables m_flat;
    std::uint64_t m_totalExposures = 0;
    std::uint64_t m_histories = 0;
};
}
