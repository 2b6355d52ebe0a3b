// xrl_lite.cpp — analytic / anchor-table photon interaction data (see xrl_lite.hpp).
//
// Models (all documented as approximate in DESIGN.md):
//  * edges / occupancies: short tables (keV) + aufbau filling with j-split subshells
//  * F(q,Z), S(q,Z): Slater-screened hydrogenic shells, closed-form Fourier transform of
//    r^(2n-2) exp(-2 zeta r) densities; S = sum_g N_g (1 - f_g^2)
//  * coherent / incoherent cross sections: numerical integration of Thomson*F^2 and
//    Klein-Nishina*S over the scattering angle, cached on a log energy grid
//  * photoelectric: oxygen anchor curve (from the well known water table) scaled with a
//    Z- and E-dependent exponent for Z<=20 (+Fe, Zn by interpolation), explicit per-branch
//    anchor points for Cu, Sn, I, W, Pb, edge jump ratios for everything below an edge
//  * fluorescence: Bambynek-type yield fits, fixed line branching ratios, line energy =
//    difference of edge energies
#include "xrl_lite.hpp"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace xrl_lite {

namespace {

    xrl_error g_error = { 1, "xrl_lite: no data" };

    inline void setError(xrl_error** e)
    {
        if (e)
            *e = &g_error;
    }

    constexpr int ZMAX = 92;
    constexpr double PI = 3.14159265358979323846;
    constexpr double AVOGADRO = 0.602214076; // 1e24 /mol  (barn*mol/g -> cm2/g factor)
    constexpr double RE2 = 0.0794078; // classical electron radius squared, barn
    constexpr double MEC2 = 510.9989461; // keV
    constexpr double KEV2ANGST = 12.398520;

    struct ElementInfo {
        int Z;
        const char* symbol;
        double A;
        double density;
        std::vector<double> edges; // keV, indexed by shell number, only occupied shells with data
    };

    // clang-format off
    const ElementInfo g_elements[] = {
        { 1, "H", 1.008, 8.375e-5, { 0.0136 } },
        { 2, "He", 4.002602, 1.663e-4, { 0.0246 } },
        { 3, "Li", 6.94, 0.534, { 0.0547, 0.0053 } },
        { 4, "Be", 9.0121831, 1.848, { 0.1115, 0.0080 } },
        { 5, "B", 10.81, 2.37, { 0.1880, 0.0126, 0.0047 } },
        { 6, "C", 12.011, 2.0, { 0.2842, 0.0180, 0.0064 } },
        { 7, "N", 14.007, 1.165e-3, { 0.4099, 0.0244, 0.0092, 0.0092 } },
        { 8, "O", 15.999, 1.332e-3, { 0.5431, 0.0285, 0.0071, 0.0071 } },
        { 9, "F", 18.998403, 1.580e-3, { 0.6967, 0.0340, 0.0086, 0.0086 } },
        { 10, "Ne", 20.1797, 8.385e-4, { 0.8702, 0.0485, 0.0217, 0.0216 } },
        { 11, "Na", 22.98976928, 0.971, { 1.0708, 0.0633, 0.0311, 0.0311, 0.0052 } },
        { 12, "Mg", 24.305, 1.74, { 1.3030, 0.0894, 0.0514, 0.0514, 0.0076 } },
        { 13, "Al", 26.9815385, 2.6989, { 1.5596, 0.1177, 0.0731, 0.0727, 0.0107, 0.0060 } },
        { 14, "Si", 28.085, 2.33, { 1.8389, 0.1487, 0.0992, 0.0992, 0.0113, 0.0051 } },
        { 15, "P", 30.973762, 2.2, { 2.1455, 0.1893, 0.1322, 0.1322, 0.0143, 0.0063, 0.0063 } },
        { 16, "S", 32.06, 2.0, { 2.4720, 0.2292, 0.1648, 0.1648, 0.0158, 0.0080, 0.0080 } },
        { 17, "Cl", 35.45, 2.995e-3, { 2.8224, 0.2702, 0.2016, 0.2000, 0.0175, 0.0068, 0.0068 } },
        { 18, "Ar", 39.948, 1.662e-3, { 3.2059, 0.3263, 0.2507, 0.2486, 0.0292, 0.0159, 0.0157 } },
        { 19, "K", 39.0983, 0.862, { 3.6084, 0.3771, 0.2963, 0.2936, 0.0339, 0.0178, 0.0178 } },
        { 20, "Ca", 40.078, 1.55, { 4.0385, 0.4378, 0.3500, 0.3464, 0.0437, 0.0254, 0.0254 } },
        { 26, "Fe", 55.845, 7.874, { 7.1120, 0.8461, 0.7211, 0.7081, 0.0929, 0.0540, 0.0540, 0.0036, 0.0036, 0.0079 } },
        { 29, "Cu", 63.546, 8.96, { 8.9789, 1.0961, 0.9510, 0.9311, 0.1198, 0.0736, 0.0736, 0.0016, 0.0016, 0.0077 } },
        { 30, "Zn", 65.38, 7.133, { 9.6586, 1.1936, 1.0428, 1.0197, 0.1359, 0.0866, 0.0866, 0.0081, 0.0081, 0.0094 } },
        { 50, "Sn", 118.71, 7.31, { 29.2001, 4.4647, 4.1561, 3.9288, 0.8838, 0.7564, 0.7144, 0.4933, 0.4848, 0.1365, 0.0886, 0.0886, 0.0239, 0.0239 } },
        { 53, "I", 126.90447, 4.93, { 33.1694, 5.1881, 4.8521, 4.5571, 1.0721, 0.9305, 0.8746, 0.6313, 0.6194, 0.1864, 0.1301, 0.1301, 0.0507, 0.0507 } },
        { 74, "W", 183.84, 19.3, { 69.5250, 12.0998, 11.5440, 10.2068, 2.8196, 2.5749, 2.2810, 1.8716, 1.8092, 0.5950, 0.4916, 0.4253, 0.2588, 0.2454, 0.0365, 0.0336, 0.0771, 0.0468, 0.0356, 0.0061, 0.0061 } },
        { 82, "Pb", 207.2, 11.35, { 88.0045, 15.8608, 15.2000, 13.0352, 3.8507, 3.5542, 3.0664, 2.5856, 2.4840, 0.8936, 0.7639, 0.6445, 0.4352, 0.4129, 0.1429, 0.1381, 0.1473, 0.1048, 0.0860, 0.0218, 0.0192 } },
    };
    // clang-format on

    const ElementInfo* findElement(int Z)
    {
        for (const auto& e : g_elements)
            if (e.Z == Z)
                return &e;
        return nullptr;
    }

    // ---- electron configuration: aufbau filling with j-split subshells -------------
    // returns occupancy per xraylib shell index
    std::array<int, N_SHELLS> configuration(int Z)
    {
        struct Orb {
            int shellLo; // first subshell index (j = l-1/2), or the s shell
            int capLo;
            int shellHi; // second subshell (j = l+1/2) or -1
            int capHi;
        };
        // filling order 1s 2s 2p 3s 3p 4s 3d 4p 5s 4d 5p 6s 4f 5d 6p
        static const Orb order[] = {
            { K_SHELL, 2, -1, 0 }, { L1_SHELL, 2, -1, 0 }, { L2_SHELL, 2, L3_SHELL, 4 },
            { M1_SHELL, 2, -1, 0 }, { M2_SHELL, 2, M3_SHELL, 4 }, { N1_SHELL, 2, -1, 0 },
            { M4_SHELL, 4, M5_SHELL, 6 }, { N2_SHELL, 2, N3_SHELL, 4 }, { O1_SHELL, 2, -1, 0 },
            { N4_SHELL, 4, N5_SHELL, 6 }, { O2_SHELL, 2, O3_SHELL, 4 }, { P1_SHELL, 2, -1, 0 },
            { N6_SHELL, 6, N7_SHELL, 8 }, { O4_SHELL, 4, O5_SHELL, 6 }, { P2_SHELL, 2, P3_SHELL, 4 }
        };
        std::array<int, N_SHELLS> occ {};
        int left = Z;
        for (const auto& o : order) {
            if (left <= 0)
                break;
            int n = std::min(left, o.capLo);
            occ[o.shellLo] += n;
            left -= n;
            if (o.shellHi >= 0 && left > 0) {
                n = std::min(left, o.capHi);
                occ[o.shellHi] += n;
                left -= n;
            }
        }
        // s -> d promotions of the elements in the table
        if (Z == 29) { // 3d10 4s1
            occ[N1_SHELL] = 1;
            occ[M4_SHELL] = 4;
            occ[M5_SHELL] = 6;
        }
        return occ;
    }

    // ---- Slater shells for F and S ---------------------------------------------------
    struct SlaterGroup {
        double nElectrons;
        double nStar;
        double a; // 2*zeta in 1/bohr
    };

    std::vector<SlaterGroup> slaterGroups(int Z)
    {
        const auto occ = configuration(Z);
        struct G {
            int n; // principal quantum number
            int kind; // 0 sp, 1 d, 2 f
            int count;
        };
        auto cnt = [&](std::initializer_list<int> shells) {
            int c = 0;
            for (int s : shells)
                c += occ[s];
            return c;
        };
        std::vector<G> g = {
            { 1, 0, cnt({ K_SHELL }) },
            { 2, 0, cnt({ L1_SHELL, L2_SHELL, L3_SHELL }) },
            { 3, 0, cnt({ M1_SHELL, M2_SHELL, M3_SHELL }) },
            { 3, 1, cnt({ M4_SHELL, M5_SHELL }) },
            { 4, 0, cnt({ N1_SHELL, N2_SHELL, N3_SHELL }) },
            { 4, 1, cnt({ N4_SHELL, N5_SHELL }) },
            { 4, 2, cnt({ N6_SHELL, N7_SHELL }) },
            { 5, 0, cnt({ O1_SHELL, O2_SHELL, O3_SHELL }) },
            { 5, 1, cnt({ O4_SHELL, O5_SHELL }) },
            { 5, 2, cnt({ O6_SHELL, O7_SHELL }) },
            { 6, 0, cnt({ P1_SHELL, P2_SHELL, P3_SHELL }) },
        };
        static const double nstar[] = { 0, 1.0, 2.0, 3.0, 3.7, 4.0, 4.2 };
        std::vector<SlaterGroup> out;
        for (std::size_t i = 0; i < g.size(); ++i) {
            if (g[i].count == 0)
                continue;
            double s = 0;
            if (g[i].kind == 0) {
                s += (g[i].n == 1 ? 0.30 : 0.35) * (g[i].count - 1);
                for (std::size_t j = 0; j < i; ++j) {
                    if (g[j].n == g[i].n - 1)
                        s += 0.85 * g[j].count;
                    else if (g[j].n < g[i].n - 1)
                        s += 1.0 * g[j].count;
                    // same n to the left cannot happen for sp groups
                }
            } else {
                s += 0.35 * (g[i].count - 1);
                for (std::size_t j = 0; j < i; ++j)
                    s += 1.0 * g[j].count;
            }
            const double zeff = std::max(Z - s, 0.5);
            const double zeta = zeff / nstar[g[i].n];
            out.push_back({ static_cast<double>(g[i].count), nstar[g[i].n], 2.0 * zeta });
        }
        return out;
    }

    // normalised form factor of one Slater shell at momentum transfer K (1/bohr)
    inline double slaterFF(const SlaterGroup& g, double K)
    {
        if (K < 1e-8)
            return 1.0;
        const double n2 = 2.0 * g.nStar;
        const double phi = std::atan(K / g.a);
        const double r2 = g.a * g.a + K * K;
        // a^(2n+1) sin(2n phi) / (2n K (a^2+K^2)^n)
        const double lg = (n2 + 1.0) * std::log(g.a) - g.nStar * std::log(r2);
        return std::exp(lg) * std::sin(n2 * phi) / (n2 * K);
    }

    struct ElementModel {
        std::vector<SlaterGroup> groups;
        // cached integrated scatter cross sections (barn/atom) on a log energy grid
        std::vector<double> lnE, lnCoh, lnIncoh;
    };

    std::mutex g_mutex;
    std::map<int, ElementModel> g_models;

    constexpr double X_TO_K = 4.0 * PI * 0.529177210903; // x[1/A] -> K[1/bohr]

    double ffFromGroups(const std::vector<SlaterGroup>& groups, double x)
    {
        const double K = X_TO_K * x;
        double F = 0;
        for (const auto& g : groups)
            F += g.nElectrons * slaterFF(g, K);
        return F;
    }
    double sfFromGroups(const std::vector<SlaterGroup>& groups, double x)
    {
        const double K = X_TO_K * x;
        double S = 0;
        for (const auto& g : groups) {
            const double f = slaterFF(g, K);
            S += g.nElectrons * (1.0 - f * f);
        }
        return S;
    }

    template <typename F>
    double integrateAngle(F integrand) // integral over mu in [-1,1] of integrand(mu, t) with t=(1-mu)/2
    {
        // t = s^3 concentrates nodes at forward angles where F^2 is peaked
        constexpr int N = 600;
        const double h = 1.0 / N;
        double sum = 0;
        for (int i = 0; i <= N; ++i) {
            const double s = i * h;
            const double t = s * s * s;
            const double w = (i == 0 || i == N) ? 1.0 : (i % 2 ? 4.0 : 2.0);
            sum += w * integrand(1.0 - 2.0 * t, t) * 3.0 * s * s;
        }
        return 2.0 * sum * h / 3.0;
    }

    // published pointers: the common path (model already built) takes no lock, so the per-material table builders
    // can run on several host threads
    std::atomic<const ElementModel*> g_modelOf[128] {};

    const ElementModel& model(int Z)
    {
        const int slot = Z & 127;
        if (const ElementModel* ready = g_modelOf[slot].load(std::memory_order_acquire))
            return *ready;
        std::lock_guard<std::mutex> lock(g_mutex);
        auto it = g_models.find(Z);
        if (it != g_models.end()) {
            g_modelOf[slot].store(&it->second, std::memory_order_release);
            return it->second;
        }
        ElementModel m;
        m.groups = slaterGroups(Z);
        constexpr int NE = 140;
        const double lmin = std::log(0.4), lmax = std::log(1300.0);
        for (int i = 0; i < NE; ++i) {
            const double le = lmin + (lmax - lmin) * i / (NE - 1);
            const double E = std::exp(le);
            const double k = E / KEV2ANGST;
            const double alpha = E / MEC2;
            const double coh = integrateAngle([&](double mu, double t) {
                const double F = ffFromGroups(m.groups, k * std::sqrt(t));
                return PI * RE2 * (1.0 + mu * mu) * F * F;
            });
            const double incoh = integrateAngle([&](double mu, double t) {
                const double r = 1.0 / (1.0 + alpha * (1.0 - mu));
                const double kn = PI * RE2 * r * r * (r + 1.0 / r - (1.0 - mu * mu));
                return kn * sfFromGroups(m.groups, k * std::sqrt(t));
            });
            m.lnE.push_back(le);
            m.lnCoh.push_back(std::log(std::max(coh, 1e-300)));
            m.lnIncoh.push_back(std::log(std::max(incoh, 1e-300)));
        }
        const ElementModel& stored = g_models.emplace(Z, std::move(m)).first->second; // std::map nodes never move
        g_modelOf[slot].store(&stored, std::memory_order_release);
        return stored;
    }

    double interpLogGrid(const std::vector<double>& x, const std::vector<double>& y, double lx)
    {
        if (lx <= x.front())
            return y.front();
        if (lx >= x.back())
            return y.back();
        const double step = (x.back() - x.front()) / (x.size() - 1);
        std::size_t i = static_cast<std::size_t>((lx - x.front()) / step);
        if (i >= x.size() - 1)
            i = x.size() - 2;
        const double f = (lx - x[i]) / (x[i + 1] - x[i]);
        return y[i] + f * (y[i + 1] - y[i]);
    }

    // ---- photoelectric model -----------------------------------------------------------
    struct Pt {
        double E, v;
    };

    // log-log interpolation with end-slope extrapolation
    double loglog(const std::vector<Pt>& p, double E)
    {
        const double le = std::log(E);
        std::size_t i = 0;
        if (E <= p.front().E)
            i = 0;
        else if (E >= p.back().E)
            i = p.size() - 2;
        else {
            while (i + 2 < p.size() && p[i + 1].E <= E)
                ++i;
        }
        const double x0 = std::log(p[i].E), x1 = std::log(p[i + 1].E);
        const double y0 = std::log(p[i].v), y1 = std::log(p[i + 1].v);
        return std::exp(y0 + (y1 - y0) * (le - x0) / (x1 - x0));
    }

    // oxygen photoelectric cross section above its K edge, cm2/g (from the water table / 0.888)
    const std::vector<Pt> g_oxygenPhoto = {
        { 0.5431, 2.35e4 }, { 1.0, 4590.0 }, { 1.5, 1549.0 }, { 2.0, 694.0 }, { 3.0, 216.0 }, { 4.0, 92.4 }, { 5.0, 47.2 },
        { 6.0, 27.15 }, { 8.0, 11.17 }, { 10.0, 5.563 }, { 15.0, 1.543 }, { 20.0, 0.6126 }, { 30.0, 0.1644 },
        { 40.0, 0.06396 }, { 50.0, 0.03097 }, { 60.0, 0.01700 }, { 80.0, 0.006531 }, { 100.0, 0.003119 },
        { 150.0, 0.000822 }, { 200.0, 0.0003254 }, { 300.0, 9.0e-5 }, { 500.0, 1.97e-5 }, { 1000.0, 3.3e-6 },
        { 1500.0, 1.4e-6 }
    };

    // exponent n in tau_atom(Z,E) = tau_atom(8,E) (Z/8)^n for three anchor elements
    const std::vector<Pt> g_nCarbon = { { 1.0, 3.54 }, { 3.0, 4.03 }, { 10.0, 4.42 }, { 30.0, 4.55 }, { 100.0, 4.75 } };
    const std::vector<Pt> g_nAluminium = { { 2.0, 3.51 }, { 3.0, 3.74 }, { 10.0, 4.21 }, { 30.0, 4.515 }, { 100.0, 4.73 } };
    const std::vector<Pt> g_nCalcium = { { 5.0, 3.78 }, { 10.0, 4.065 }, { 30.0, 4.406 }, { 100.0, 4.647 } };

    double linLogE(const std::vector<Pt>& p, double E) // linear in ln E, end-slope extrapolation (clamped range)
    {
        E = std::clamp(E, 0.3, 2000.0);
        const double le = std::log(E);
        std::size_t i = 0;
        if (E <= p.front().E)
            i = 0;
        else if (E >= p.back().E)
            i = p.size() - 2;
        else {
            while (i + 2 < p.size() && p[i + 1].E <= E)
                ++i;
        }
        const double x0 = std::log(p[i].E), x1 = std::log(p[i + 1].E);
        return p[i].v + (p[i + 1].v - p[i].v) * (le - x0) / (x1 - x0);
    }

    // K-branch (all shells active) photo cross section in barn/atom for Z <= 20
    double lightKBranchBarn(int Z, double E)
    {
        const double tauO = loglog(g_oxygenPhoto, E) * 15.999 / AVOGADRO; // barn/atom
        if (Z == 8)
            return tauO;
        const double n6 = linLogE(g_nCarbon, E);
        const double n13 = linLogE(g_nAluminium, E);
        const double n20 = linLogE(g_nCalcium, E);
        const double lz = std::log(static_cast<double>(Z));
        const double l6 = std::log(6.0), l13 = std::log(13.0), l20 = std::log(20.0);
        double n;
        if (Z <= 13)
            n = n6 + (n13 - n6) * (lz - l6) / (l13 - l6);
        else
            n = n13 + (n20 - n13) * (lz - l13) / (l20 - l13);
        return tauO * std::pow(Z / 8.0, n);
    }

    struct Branch {
        double Emin, Emax;
        std::vector<Pt> pts; // cm2/g
    };

    // clang-format off
    const std::map<int, std::vector<Branch>> g_heavyBranches = {
        { 29, {
            { 8.9789, 1e9, { { 8.979, 277.0 }, { 10, 214.0 }, { 15, 73.2 }, { 20, 33.0 }, { 30, 10.45 }, { 40, 4.51 }, { 50, 2.33 }, { 60, 1.36 }, { 80, 0.565 }, { 100, 0.285 }, { 150, 0.0832 }, { 200, 0.0359 }, { 300, 0.0112 }, { 500, 0.0030 }, { 1000, 0.00060 } } },
            { 1.0961, 8.9789, { { 1.0961, 9300.0 }, { 1.5, 4410.0 }, { 2, 2150.0 }, { 3, 746.0 }, { 4, 345.0 }, { 5, 188.5 }, { 6, 114.5 }, { 8, 51.3 }, { 8.979, 37.0 } } } } },
        { 50, {
            { 29.2001, 1e9, { { 29.2, 43.0 }, { 30, 40.6 }, { 40, 18.9 }, { 50, 10.3 }, { 60, 6.2 }, { 80, 2.77 }, { 100, 1.47 }, { 150, 0.46 }, { 200, 0.205 }, { 300, 0.066 }, { 500, 0.0175 }, { 1000, 0.0035 } } },
            { 4.4647, 29.2001, { { 4.465, 1135.0 }, { 5, 855.0 }, { 6, 540.0 }, { 8, 250.0 }, { 10, 137.5 }, { 15, 45.0 }, { 20, 20.7 }, { 29.2, 7.1 } } } } },
        { 53, {
            { 33.1694, 1e9, { { 33.17, 35.2 }, { 40, 22.4 }, { 50, 11.9 }, { 60, 7.2 }, { 80, 3.24 }, { 100, 1.72 }, { 150, 0.54 }, { 200, 0.242 }, { 300, 0.078 }, { 500, 0.0205 }, { 1000, 0.0041 } } },
            { 5.1881, 33.1694, { { 5.188, 900.0 }, { 6, 620.0 }, { 8, 292.0 }, { 10, 160.0 }, { 15, 54.0 }, { 20, 24.6 }, { 30, 8.0 }, { 33.17, 6.2 } } } } },
        { 74, {
            { 69.5250, 1e9, { { 69.525, 10.9 }, { 80, 7.5 }, { 100, 4.21 }, { 150, 1.44 }, { 200, 0.67 }, { 300, 0.226 }, { 500, 0.062 }, { 1000, 0.0125 } } },
            { 12.0998, 69.5250, { { 12.1, 240.0 }, { 15, 137.0 }, { 20, 64.3 }, { 30, 22.0 }, { 40, 10.15 }, { 50, 5.53 }, { 60, 3.35 }, { 69.525, 2.2 } } },
            { 2.8196, 10.2068, { { 3, 1900.0 }, { 4, 980.0 }, { 5, 565.0 }, { 6, 359.0 }, { 8, 171.0 }, { 10.2, 92.0 } } } } },
        { 82, {
            { 88.0045, 1e9, { { 88.0, 7.4 }, { 100, 5.34 }, { 150, 1.86 }, { 200, 0.83 }, { 300, 0.283 }, { 500, 0.079 }, { 1000, 0.0165 } } },
            { 15.8608, 88.0045, { { 15.86, 153.0 }, { 20, 84.9 }, { 30, 29.5 }, { 40, 13.8 }, { 50, 7.6 }, { 60, 4.65 }, { 80, 2.12 }, { 88.0, 1.64 } } },
            { 3.8507, 13.0352, { { 3.85, 1360.0 }, { 4, 1248.0 }, { 5, 728.0 }, { 6, 465.0 }, { 8, 226.0 }, { 10, 129.0 }, { 13.035, 66.0 } } } } },
    };
    // clang-format on

    double jumpRatio(int Z, int shell)
    {
        switch (shell) {
        case K_SHELL:
            return 10.9 * std::pow(13.0 / Z, 0.48);
        case L1_SHELL:
            return 1.16;
        case L2_SHELL:
            return 1.40;
        case L3_SHELL:
            return std::max(3.9 - 0.018 * Z, 2.0);
        case M1_SHELL:
            return 1.04;
        case M2_SHELL:
            return 1.06;
        case M3_SHELL:
            return 1.14;
        case M4_SHELL:
            return 1.30;
        case M5_SHELL:
            return 1.50;
        default:
            return 1.02;
        }
    }

    // smooth "all shells active" K-branch, barn/atom, for every supported Z
    double kBranchBarn(int Z, double E)
    {
        if (Z <= 20)
            return lightKBranchBarn(Z, E);
        auto heavy = g_heavyBranches.find(Z);
        if (heavy != g_heavyBranches.end()) {
            const auto* el = findElement(Z);
            return loglog(heavy->second.front().pts, E) * el->A / AVOGADRO;
        }
        // Fe, Zn: interpolate ln(tau_atom) in ln Z between Ca and Cu K-branches
        const double t20 = std::log(lightKBranchBarn(20, E));
        const double t29 = std::log(loglog(g_heavyBranches.at(29).front().pts, E) * 63.546 / AVOGADRO);
        const double f = (std::log(static_cast<double>(Z)) - std::log(20.0)) / (std::log(29.0) - std::log(20.0));
        return std::exp(t20 + f * (t29 - t20));
    }

    // total photoelectric cross section, barn/atom
    double photoBarn(int Z, double E)
    {
        const auto* el = findElement(Z);
        if (!el || E <= 0)
            return 0;
        auto heavy = g_heavyBranches.find(Z);
        if (heavy != g_heavyBranches.end()) {
            const auto& br = heavy->second;
            for (const auto& b : br)
                if (E >= b.Emin && E < b.Emax)
                    return loglog(b.pts, E) * el->A / AVOGADRO;
            // not inside an explicit branch: extrapolate the nearest branch above and divide by the jumps crossed
            const Branch* above = nullptr;
            for (const auto& b : br)
                if (b.Emin > E && (!above || b.Emin < above->Emin))
                    above = &b;
            double v = loglog(above->pts, E) * el->A / AVOGADRO;
            for (std::size_t s = 0; s < el->edges.size(); ++s)
                if (el->edges[s] > E && el->edges[s] <= above->Emin * (1 + 1e-9))
                    v /= jumpRatio(Z, static_cast<int>(s));
            return v;
        }
        double v = kBranchBarn(Z, E);
        for (std::size_t s = 0; s < el->edges.size(); ++s)
            if (el->edges[s] > E)
                v /= jumpRatio(Z, static_cast<int>(s));
        return v;
    }

    // Klein-Nishina energy-transfer fraction for mu_en
    double knTransferFraction(double E)
    {
        const double alpha = E / MEC2;
        double num = 0, den = 0;
        constexpr int N = 200;
        for (int i = 0; i <= N; ++i) {
            const double mu = -1.0 + 2.0 * i / N;
            const double w = (i == 0 || i == N) ? 1.0 : (i % 2 ? 4.0 : 2.0);
            const double r = 1.0 / (1.0 + alpha * (1.0 - mu));
            const double kn = r * r * (r + 1.0 / r - (1.0 - mu * mu));
            num += w * kn * (1.0 - r);
            den += w * kn;
        }
        return num / den;
    }

    // ---- compounds -----------------------------------------------------------------------
    struct Compound {
        std::string name;
        std::vector<int> Z;
        std::vector<double> w; // mass fractions
        std::vector<double> nAtoms;
        double density = -1;
    };

    // clang-format off
    const std::vector<Compound> g_nist = {
        { "A-150 Tissue-Equivalent Plastic", { 1, 6, 7, 8, 9, 20 }, { 0.101327, 0.775501, 0.035057, 0.052316, 0.017422, 0.018378 }, {}, 1.127 },
        { "Adipose Tissue (ICRP)", { 1, 6, 7, 8, 11, 12, 15, 16, 17, 19, 20, 26, 30 }, { 0.119477, 0.637240, 0.007970, 0.232333, 0.000500, 0.000020, 0.000160, 0.000730, 0.001190, 0.000320, 0.000020, 0.000020, 0.000020 }, {}, 0.92 },
        { "Air, Dry (near sea level)", { 6, 7, 8, 18 }, { 0.000124, 0.755267, 0.231781, 0.012827 }, {}, 0.00120479 },
        { "Aluminum Oxide", { 8, 13 }, { 0.470749, 0.529251 }, {}, 3.97 },
        { "Blood (ICRP)", { 1, 6, 7, 8, 11, 12, 14, 15, 16, 17, 19, 20, 26, 30 }, { 0.101866, 0.100020, 0.029640, 0.759414, 0.001850, 0.000040, 0.000030, 0.000350, 0.001850, 0.002780, 0.001630, 0.000060, 0.000460, 0.000010 }, {}, 1.06 },
        { "Bone, Compact (ICRU)", { 1, 6, 7, 8, 12, 15, 16, 20 }, { 0.063984, 0.278000, 0.027000, 0.410016, 0.002000, 0.070000, 0.002000, 0.147000 }, {}, 1.85 },
        { "Bone, Cortical (ICRP)", { 1, 6, 7, 8, 12, 15, 16, 20, 30 }, { 0.047234, 0.144330, 0.041990, 0.446096, 0.002200, 0.104970, 0.003150, 0.209930, 0.000100 }, {}, 1.85 },
        { "Brain (ICRP)", { 1, 6, 7, 8, 11, 12, 15, 16, 17, 19, 20, 26, 30 }, { 0.110667, 0.125420, 0.013280, 0.737723, 0.001840, 0.000150, 0.003540, 0.001770, 0.002360, 0.003100, 0.000090, 0.000050, 0.000010 }, {}, 1.03 },
        { "Lung (ICRP)", { 1, 6, 7, 8, 11, 12, 15, 16, 17, 19, 20, 26, 30 }, { 0.101278, 0.102310, 0.028650, 0.757072, 0.001840, 0.000730, 0.000800, 0.002250, 0.002660, 0.001940, 0.000090, 0.000370, 0.000010 }, {}, 1.05 },
        { "Muscle, Skeletal", { 1, 6, 7, 8, 11, 12, 15, 16, 17, 19, 20, 26, 30 }, { 0.100637, 0.107830, 0.027680, 0.754773, 0.000750, 0.000190, 0.001800, 0.002410, 0.000790, 0.003020, 0.000030, 0.000040, 0.000050 }, {}, 1.04 },
        { "Muscle, Striated", { 1, 6, 7, 8, 11, 12, 15, 16, 19 }, { 0.101997, 0.123000, 0.035000, 0.729003, 0.000800, 0.000200, 0.002000, 0.005000, 0.003000 }, {}, 1.04 },
        { "Polycarbonate (Makrolon, Lexan)", { 1, 6, 8 }, { 0.055491, 0.755751, 0.188758 }, {}, 1.2 },
        { "Polyethylene", { 1, 6 }, { 0.143711, 0.856289 }, {}, 0.94 },
        { "Polymethyl Methacralate (Lucite, Perspex)", { 1, 6, 8 }, { 0.080538, 0.599848, 0.319614 }, {}, 1.19 },
        { "Polystyrene", { 1, 6 }, { 0.077418, 0.922582 }, {}, 1.06 },
        { "Polytetrafluoroethylene (Teflon)", { 6, 9 }, { 0.240183, 0.759817 }, {}, 2.2 },
        { "Skin (ICRP)", { 1, 6, 7, 8, 11, 12, 15, 16, 17, 19, 20, 26, 30 }, { 0.100588, 0.228250, 0.046420, 0.619002, 0.000070, 0.000060, 0.000330, 0.001590, 0.002670, 0.000850, 0.000150, 0.000010, 0.000010 }, {}, 1.1 },
        { "Tissue, Soft (ICRP)", { 1, 6, 7, 8, 11, 12, 15, 16, 17, 19, 20, 26, 30 }, { 0.104472, 0.232190, 0.024880, 0.630238, 0.001130, 0.000130, 0.001330, 0.001990, 0.001340, 0.001990, 0.000230, 0.000050, 0.000030 }, {}, 1.0 },
        { "Tissue, Soft (ICRU four-component)", { 1, 6, 7, 8 }, { 0.101172, 0.111000, 0.026000, 0.761828 }, {}, 1.0 },
        { "Urea", { 1, 6, 7, 8 }, { 0.067131, 0.199999, 0.466459, 0.266411 }, {}, 1.323 },
        { "Water, Liquid", { 1, 8 }, { 0.111894, 0.888106 }, {}, 1.0 },
    };
    // clang-format on

    const Compound* findNIST(const char* name)
    {
        if (!name)
            return nullptr;
        for (const auto& c : g_nist)
            if (c.name == name)
                return &c;
        return nullptr;
    }

    // recursive descent parser of chemical formulas with fractional counts and parentheses
    struct FormulaParser {
        const char* p;
        bool ok = true;
        std::map<int, double> parseGroup(bool nested)
        {
            std::map<int, double> atoms;
            while (*p && ok) {
                if (*p == '(') {
                    ++p;
                    auto inner = parseGroup(true);
                    if (*p != ')') {
                        ok = false;
                        break;
                    }
                    ++p;
                    const double n = parseNumber();
                    for (auto& [z, c] : inner)
                        atoms[z] += c * n;
                } else if (*p == ')') {
                    if (!nested)
                        ok = false;
                    break;
                } else if (*p >= 'A' && *p <= 'Z') {
                    std::string sym(1, *p++);
                    while (*p >= 'a' && *p <= 'z')
                        sym.push_back(*p++);
                    const int z = SymbolToAtomicNumber(sym.c_str(), nullptr);
                    if (z <= 0) {
                        ok = false;
                        break;
                    }
                    atoms[z] += parseNumber();
                } else {
                    ok = false;
                }
            }
            return atoms;
        }
        double parseNumber()
        {
            const char* s = p;
            while ((*p >= '0' && *p <= '9') || *p == '.')
                ++p;
            if (p == s)
                return 1.0;
            return std::strtod(std::string(s, p).c_str(), nullptr);
        }
    };

    bool parseFormula(const char* str, Compound& out)
    {
        if (!str || !*str)
            return false;
        FormulaParser fp { str };
        auto atoms = fp.parseGroup(false);
        if (!fp.ok || *fp.p != 0 || atoms.empty())
            return false;
        double mass = 0;
        for (auto& [z, n] : atoms) {
            if (!elementSupported(z) || n <= 0)
                return false;
            mass += n * findElement(z)->A;
        }
        out.name = str;
        for (auto& [z, n] : atoms) {
            out.Z.push_back(z);
            out.nAtoms.push_back(n);
            out.w.push_back(n * findElement(z)->A / mass);
        }
        return true;
    }

    std::mutex g_compoundMutex;
    std::map<std::string, Compound> g_compoundCache;

    // resolves a NIST name first, then a formula (same precedence as Material's constructor)
    const Compound* resolveCompound(const char* s)
    {
        if (const Compound* n = findNIST(s))
            return n;
        std::lock_guard<std::mutex> lock(g_compoundMutex);
        auto it = g_compoundCache.find(s ? s : "");
        if (it != g_compoundCache.end())
            return &it->second;
        Compound c;
        if (!parseFormula(s, c))
            return nullptr;
        return &g_compoundCache.emplace(s, std::move(c)).first->second;
    }

    template <typename F>
    double compoundSum(const char* compound, xrl_error** error, F f)
    {
        const Compound* c = resolveCompound(compound);
        if (!c) {
            setError(error);
            return 0;
        }
        double s = 0;
        for (std::size_t i = 0; i < c->Z.size(); ++i)
            s += c->w[i] * f(c->Z[i]);
        return s;
    }

    // shell index of upper/lower level of an emission line (xraylib line numbering)
    bool lineShells(int line, int& from, int& to)
    {
        // K lines -1..-29 : KL1 KL2 KL3 KM1..KM5 KN1..KN7 KO1..KO7 KP1..KP5
        if (line <= -1 && line >= -29) {
            from = K_SHELL;
            to = -line; // -1 -> L1 (1) ... -27 -> P5 (27)
            return to < N_SHELLS + 2;
        }
        // L1 lines -30..-58 : L1L2 L1L3 L1M1..
        if (line <= -30 && line >= -58) {
            from = L1_SHELL;
            to = L2_SHELL + (-30 - line);
            return true;
        }
        // L2 lines -59..-85 : L2L3 L2M1 ..
        if (line <= -59 && line >= -85) {
            from = L2_SHELL;
            to = L3_SHELL + (-59 - line);
            return true;
        }
        // L3 lines -86..-113 : L3M1 L3M2 ...
        if (line <= -86 && line >= -113) {
            from = L3_SHELL;
            to = M1_SHELL + (-86 - line);
            return true;
        }
        return false;
    }

} // namespace

bool elementSupported(int Z) { return findElement(Z) != nullptr; }

double AtomicWeight(int Z, xrl_error** error)
{
    if (const auto* e = findElement(Z))
        return e->A;
    setError(error);
    return 0;
}

double ElementDensity(int Z, xrl_error** error)
{
    if (const auto* e = findElement(Z))
        return e->density;
    setError(error);
    return 0;
}

char* AtomicNumberToSymbol(int Z, xrl_error** error)
{
    const auto* e = findElement(Z);
    if (!e) {
        setError(error);
        return nullptr;
    }
    char* s = static_cast<char*>(std::malloc(std::strlen(e->symbol) + 1));
    std::strcpy(s, e->symbol);
    return s;
}

int SymbolToAtomicNumber(const char* symbol, xrl_error** error)
{
    if (symbol)
        for (const auto& e : g_elements)
            if (std::strcmp(e.symbol, symbol) == 0)
                return e.Z;
    setError(error);
    return 0;
}

void xrlFree(void* p) { std::free(p); }

double CS_Photo(int Z, double E, xrl_error** error)
{
    const auto* e = findElement(Z);
    if (!e || E <= 0) {
        setError(error);
        return 0;
    }
    return photoBarn(Z, E) * AVOGADRO / e->A;
}

double CS_Rayl(int Z, double E, xrl_error** error)
{
    const auto* e = findElement(Z);
    if (!e || E <= 0) {
        setError(error);
        return 0;
    }
    const auto& m = model(Z);
    return std::exp(interpLogGrid(m.lnE, m.lnCoh, std::log(E))) * AVOGADRO / e->A;
}

double CS_Compt(int Z, double E, xrl_error** error)
{
    const auto* e = findElement(Z);
    if (!e || E <= 0) {
        setError(error);
        return 0;
    }
    const auto& m = model(Z);
    return std::exp(interpLogGrid(m.lnE, m.lnIncoh, std::log(E))) * AVOGADRO / e->A;
}

double CS_Total(int Z, double E, xrl_error** error)
{
    if (!findElement(Z) || E <= 0) {
        setError(error);
        return 0;
    }
    return CS_Photo(Z, E, nullptr) + CS_Compt(Z, E, nullptr) + CS_Rayl(Z, E, nullptr);
}

double CS_Energy(int Z, double E, xrl_error** error)
{
    if (!findElement(Z) || E <= 0) {
        setError(error);
        return 0;
    }
    return CS_Photo(Z, E, nullptr) + CS_Compt(Z, E, nullptr) * knTransferFraction(E);
}

double CS_Total_CP(const char* c, double E, xrl_error** error)
{
    return compoundSum(c, error, [=](int Z) { return CS_Total(Z, E, nullptr); });
}
double CS_Photo_CP(const char* c, double E, xrl_error** error)
{
    return compoundSum(c, error, [=](int Z) { return CS_Photo(Z, E, nullptr); });
}
double CS_Rayl_CP(const char* c, double E, xrl_error** error)
{
    return compoundSum(c, error, [=](int Z) { return CS_Rayl(Z, E, nullptr); });
}
double CS_Compt_CP(const char* c, double E, xrl_error** error)
{
    return compoundSum(c, error, [=](int Z) { return CS_Compt(Z, E, nullptr); });
}
double CS_Energy_CP(const char* c, double E, xrl_error** error)
{
    return compoundSum(c, error, [=](int Z) { return CS_Energy(Z, E, nullptr); });
}

double FF_Rayl(int Z, double q, xrl_error** error)
{
    if (!findElement(Z) || q < 0) {
        setError(error);
        return 0;
    }
    return ffFromGroups(model(Z).groups, q);
}

double SF_Compt(int Z, double q, xrl_error** error)
{
    if (!findElement(Z) || q < 0) {
        setError(error);
        return 0;
    }
    return sfFromGroups(model(Z).groups, q);
}

double EdgeEnergy(int Z, int shell, xrl_error** error)
{
    const auto* e = findElement(Z);
    if (!e || shell < 0 || static_cast<std::size_t>(shell) >= e->edges.size() || e->edges[shell] <= 0) {
        setError(error);
        return 0;
    }
    return e->edges[shell];
}

double ElectronConfig(int Z, int shell, xrl_error** error)
{
    if (!findElement(Z) || shell < 0 || shell >= N_SHELLS) {
        setError(error);
        return 0;
    }
    const auto occ = configuration(Z);
    if (occ[shell] == 0) {
        setError(error);
        return 0;
    }
    return occ[shell];
}

double ComptonProfile_Partial(int Z, int shell, double pz, xrl_error** error)
{
    xrl_error* err = nullptr;
    const double eb = EdgeEnergy(Z, shell, &err);
    if (err || pz < 0) {
        setError(error);
        return 0;
    }
    // hydrogenic 1s-like profile J(pz) = 8 p0^5 / (3 pi (p0^2 + pz^2)^3), p0 = sqrt(Eb/Ry) in a.u.
    const double p0 = std::sqrt(std::max(eb, 1e-4) / 0.0136057);
    const double d = p0 * p0 + pz * pz;
    return 8.0 * std::pow(p0, 5) / (3.0 * PI * d * d * d);
}

double FluorYield(int Z, int shell, xrl_error** error)
{
    xrl_error* err = nullptr;
    EdgeEnergy(Z, shell, &err);
    if (err || shell > L3_SHELL) {
        setError(error);
        return 0;
    }
    if (shell == K_SHELL) {
        const double x = 0.015 + 0.0327 * Z - 0.64e-6 * Z * Z * Z;
        const double x4 = x * x * x * x;
        return x4 / (1.0 + x4);
    }
    const double z4 = std::pow(static_cast<double>(Z), 4);
    const double wl3 = z4 / (z4 + 1.02e8);
    if (shell == L3_SHELL)
        return wl3;
    if (shell == L2_SHELL)
        return std::min(1.1 * wl3, 1.0);
    return 0.6 * wl3;
}

double CosKronTransProb(int Z, int trans, xrl_error** error)
{
    xrl_error* err = nullptr;
    EdgeEnergy(Z, L3_SHELL, &err);
    if (err) {
        setError(error);
        return 0;
    }
    switch (trans) {
    case FL12_TRANS:
        return 0.10;
    case FL13_TRANS:
        return 0.30;
    case FL23_TRANS:
        return 0.12;
    default:
        setError(error);
        return 0;
    }
}

double LineEnergy(int Z, int line, xrl_error** error)
{
    int from, to;
    if (!lineShells(line, from, to)) {
        setError(error);
        return 0;
    }
    xrl_error* e1 = nullptr;
    xrl_error* e2 = nullptr;
    const double a = EdgeEnergy(Z, from, &e1);
    const double b = EdgeEnergy(Z, to, &e2);
    if (e1 || e2 || a <= b) {
        setError(error);
        return 0;
    }
    return a - b;
}

double RadRate(int Z, int line, xrl_error** error)
{
    int from, to;
    xrl_error* err = nullptr;
    if (!lineShells(line, from, to) || LineEnergy(Z, line, &err) <= 0 || err) {
        setError(error);
        return 0;
    }
    xrl_error* occErr = nullptr;
    ElectronConfig(Z, to, &occErr);
    if (occErr) {
        setError(error);
        return 0;
    }
    // dipole-allowed lines with typical branching ratios; the K-beta share grows with Z
    const double kb = std::clamp(0.004 * (Z - 10), 0.0, 0.22); // total K-beta fraction
    double r = 0;
    if (from == K_SHELL) {
        if (to == L3_SHELL)
            r = (1 - kb) * 0.66;
        else if (to == L2_SHELL)
            r = (1 - kb) * 0.34;
        else if (to == M3_SHELL)
            r = kb * 0.55;
        else if (to == M2_SHELL)
            r = kb * 0.28;
        else if (to == N3_SHELL)
            r = kb * 0.11;
        else if (to == N2_SHELL)
            r = kb * 0.06;
    } else if (from == L1_SHELL) {
        if (to == M3_SHELL)
            r = 0.45;
        else if (to == M2_SHELL)
            r = 0.33;
        else if (to == N3_SHELL)
            r = 0.12;
        else if (to == N2_SHELL)
            r = 0.10;
    } else if (from == L2_SHELL) {
        if (to == M4_SHELL)
            r = 0.80;
        else if (to == N4_SHELL)
            r = 0.15;
        else if (to == M1_SHELL)
            r = 0.05;
    } else if (from == L3_SHELL) {
        if (to == M5_SHELL)
            r = 0.72;
        else if (to == M4_SHELL)
            r = 0.08;
        else if (to == N5_SHELL)
            r = 0.16;
        else if (to == M1_SHELL)
            r = 0.04;
    }
    if (r <= 0) {
        setError(error);
        return 0;
    }
    return r;
}

double CSb_Photo_Partial(int Z, int shell, double E, xrl_error** error)
{
    const auto* el = findElement(Z);
    xrl_error* err = nullptr;
    const double edge = EdgeEnergy(Z, shell, &err);
    if (!el || err || E < edge || E <= 0) {
        setError(error);
        return 0;
    }
    // distribute the total over the shells that can be ionised, from the deepest outwards
    std::vector<int> order(el->edges.size());
    for (std::size_t i = 0; i < order.size(); ++i)
        order[i] = static_cast<int>(i);
    std::sort(order.begin(), order.end(), [&](int a, int b) { return el->edges[a] > el->edges[b]; });
    const double total = photoBarn(Z, E);
    double remaining = 1.0;
    int lastActive = -1;
    for (int s : order)
        if (E >= el->edges[s])
            lastActive = s;
    for (int s : order) {
        if (E < el->edges[s])
            continue;
        const double r = jumpRatio(Z, s);
        const double share = (s == lastActive) ? remaining : remaining * (1.0 - 1.0 / r);
        if (s == shell)
            return total * share;
        remaining /= r;
    }
    setError(error);
    return 0;
}

compoundData* CompoundParser(const char* compound, xrl_error** error)
{
    Compound c;
    if (!parseFormula(compound, c)) {
        setError(error);
        return nullptr;
    }
    auto* cd = static_cast<compoundData*>(std::malloc(sizeof(compoundData)));
    cd->nElements = static_cast<int>(c.Z.size());
    cd->Elements = static_cast<int*>(std::malloc(sizeof(int) * c.Z.size()));
    cd->massFractions = static_cast<double*>(std::malloc(sizeof(double) * c.Z.size()));
    cd->nAtoms = static_cast<double*>(std::malloc(sizeof(double) * c.Z.size()));
    cd->nAtomsAll = 0;
    cd->molarMass = 0;
    for (std::size_t i = 0; i < c.Z.size(); ++i) {
        cd->Elements[i] = c.Z[i];
        cd->massFractions[i] = c.w[i];
        cd->nAtoms[i] = c.nAtoms[i];
        cd->nAtomsAll += c.nAtoms[i];
        cd->molarMass += c.nAtoms[i] * findElement(c.Z[i])->A;
    }
    return cd;
}

void FreeCompoundData(compoundData* cd)
{
    if (!cd)
        return;
    std::free(cd->Elements);
    std::free(cd->massFractions);
    std::free(cd->nAtoms);
    std::free(cd);
}

compoundDataNIST* GetCompoundDataNISTByName(const char* name, xrl_error** error)
{
    const Compound* c = findNIST(name);
    if (!c) {
        setError(error);
        return nullptr;
    }
    auto* cd = static_cast<compoundDataNIST*>(std::malloc(sizeof(compoundDataNIST)));
    cd->name = static_cast<char*>(std::malloc(c->name.size() + 1));
    std::strcpy(cd->name, c->name.c_str());
    cd->nElements = static_cast<int>(c->Z.size());
    cd->Elements = static_cast<int*>(std::malloc(sizeof(int) * c->Z.size()));
    cd->massFractions = static_cast<double*>(std::malloc(sizeof(double) * c->Z.size()));
    for (std::size_t i = 0; i < c->Z.size(); ++i) {
        cd->Elements[i] = c->Z[i];
        cd->massFractions[i] = c->w[i];
    }
    cd->density = c->density;
    return cd;
}

void FreeCompoundDataNIST(compoundDataNIST* cd)
{
    if (!cd)
        return;
    std::free(cd->name);
    std::free(cd->Elements);
    std::free(cd->massFractions);
    std::free(cd);
}

char** GetCompoundDataNISTList(int* nCompounds, xrl_error**)
{
    char** list = static_cast<char**>(std::malloc(sizeof(char*) * g_nist.size()));
    for (std::size_t i = 0; i < g_nist.size(); ++i) {
        list[i] = static_cast<char*>(std::malloc(g_nist[i].name.size() + 1));
        std::strcpy(list[i], g_nist[i].name.c_str());
    }
    if (nCompounds)
        *nCompounds = static_cast<int>(g_nist.size());
    return list;
}

} // namespace xrl_lite
