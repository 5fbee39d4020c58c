// subrecon_b200_host.hpp — C++ host-side mirror of the reference's interface around the hot path.
//
// The reference is Java and this image has no JVM, so the host side above the C ABI is C++ (header-only, C++17).
// Names, argument meaning and error behaviour follow the Java classes the native path plugs into:
//
//   subrecon::SiteResult                 /root/reference/src/subrecon/recon/SiteResult.java:52-139
//   subrecon::JointBranchReconstruction  /root/reference/src/subrecon/recon/JointBranchReconstruction.java:49-80
//   subrecon::Utils::roundDouble         /root/reference/src/subrecon/utils/Utils.java:28-31
//   subrecon::runColumns                 the ordered join + print loop, /root/reference/src/subrecon/SubRecon.java:122-150
//
// One JointBranchReconstruction object covers a BATCH of columns (call() returns one SiteResult per column): the
// per-column Callable + thread pool is exactly what libsubrecon_b200.so replaces. All arithmetic of the path happens
// in the library; this header only flattens inputs and formats outputs. No CPU fallback: the constructor throws if
// sr_create fails.
#ifndef SUBRECON_B200_HOST_HPP
#define SUBRECON_B200_HOST_HPP

#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "subrecon_b200.h"

namespace subrecon {

struct Constants {                                              // Constants.java:26-39
    static constexpr const char* DELIM = "\t";
    static constexpr const char* SUB_PROB_DELIM = ":";
    static constexpr double DEFAULT_PRINT_THRESHOLD = 0.5;
    static constexpr int DEFAULT_SIG_DIGITS = 2;
    static constexpr double EPSILON = 1e-6;
    static constexpr const char* AA_ORDER = "ARNDCQEGHILKMFPSTWYV";   // pal.datatype.AminoAcids state order
};

struct Utils {
    // (double) Math.round(x * (long)10^d) / (long)10^d, Math.round = floor(x + 1/2)
    static double roundDouble(double toRound, int decPlaces) {
        const long long ten = (long long)std::pow(10.0, (double)decPlaces);
        const double y = toRound * (double)ten;
        long long r;
        if (y != y) r = 0;
        else if (y >= 9.2233720368547758e18) r = INT64_MAX;
        else if (y <= -9.2233720368547758e18) r = INT64_MIN;
        else r = (long long)std::floor(y + 0.5);
        return (double)r / (double)ten;
    }
};

// java.lang.Double.toString: shortest digits that round-trip; plain decimal for 1e-3 <= |x| < 1e7, else d.dddE±n
inline std::string javaDoubleToString(double x) {
    if (x != x) return "NaN";
    if (std::isinf(x)) return x > 0 ? "Infinity" : "-Infinity";
    if (x == 0.0) return std::signbit(x) ? "-0.0" : "0.0";
    const double a = std::fabs(x);
    char buf[64];
    auto res = std::to_chars(buf, buf + sizeof(buf), a, std::chars_format::scientific);   // shortest round-trip
    std::string sci(buf, res.ptr);                                                        // d[.ddd]e±XX
    const size_t epos = sci.find('e');
    std::string mant = sci.substr(0, epos);
    const int exp10 = std::stoi(sci.substr(epos + 1));
    std::string digits;
    for (char c : mant) if (c != '.') digits.push_back(c);
    std::string out;
    if (a >= 1e-3 && a < 1e7) {
        if (exp10 >= 0) {
            std::string ip = digits.substr(0, std::min<size_t>(digits.size(), (size_t)exp10 + 1));
            while ((int)ip.size() < exp10 + 1) ip.push_back('0');
            std::string fp = digits.size() > (size_t)exp10 + 1 ? digits.substr((size_t)exp10 + 1) : "0";
            out = ip + "." + fp;
        } else {
            out = "0." + std::string((size_t)(-exp10 - 1), '0') + digits;
        }
    } else {
        out = digits.substr(0, 1) + "." + (digits.size() > 1 ? digits.substr(1) : "0") + "E" + std::to_string(exp10);
    }
    return x < 0 ? "-" + out : out;
}

class SiteResult {
public:
    // SiteResult(int site, double marginalLnL, double[][] branchProbs, double threshold, boolean sortByProb, int sigDigits)
    SiteResult(int site, double marginalLnL, const double* branchProbs /* [20][20], row = state at node A */,
               double threshold = Constants::DEFAULT_PRINT_THRESHOLD, bool sortByProb = true,
               int sigDigits = Constants::DEFAULT_SIG_DIGITS)
        : site_(site), marginalLnL_(marginalLnL), maxIIProb_(-1.0), maxProb_(-1.0), sigDigits_(sigDigits) {
        for (int i = 0; i < 20; ++i)
            for (int j = 0; j < 20; ++j) {                                    // SiteResult.java:69-83
                const double p = branchProbs[i * 20 + j];
                if (i == j) maxIIProb_ = std::fmax(maxIIProb_, p);
                maxProb_ = std::fmax(maxProb_, p);
                if (p >= threshold) {
                    aboveThreshProbs_.push_back(p);
                    aboveThreshSubs_.push_back(std::string{Constants::AA_ORDER[i], Constants::AA_ORDER[j]});
                }
            }
        if (sortByProb) sortByProbImpl();
    }
    // From one compact record (device-side scan, sr_compact layout of capacity `cap`), the way the Java host does it
    // (java/subrecon/recon/NativeRecon.java::reconstruct): through the constructor above, on a matrix that holds the kept
    // entries plus carriers for maxIIProb / maxProb when those are below the threshold.
    static SiteResult fromCompact(int site, const unsigned char* rec, int cap, double threshold, bool sortByProb = true,
                                  int sigDigits = Constants::DEFAULT_SIG_DIGITS) {
        double lnl, maxII, maxP;
        int32_t kept;
        std::memcpy(&lnl, rec, 8); std::memcpy(&maxII, rec + 8, 8); std::memcpy(&maxP, rec + 16, 8); std::memcpy(&kept, rec + 24, 4);
        if (kept > cap) throw std::runtime_error("compact record overflow");
        const unsigned char* pairs = rec + 32;
        const unsigned char* probs = rec + 32 + (2 * cap + 7) / 8 * 8;
        double m[400] = {0.0};
        for (int k = 0; k < kept; ++k) {
            int16_t ab; double pr;
            std::memcpy(&ab, pairs + 2 * k, 2); std::memcpy(&pr, probs + 8 * k, 8);
            m[ab] = pr;
        }
        if (maxII < threshold) m[0] = std::fmax(m[0], maxII);
        if (maxP < threshold && maxP > maxII) m[1] = std::fmax(m[1], maxP);
        return SiteResult(site, lnl, m, threshold, sortByProb, sigDigits);
    }
    int getSite() const { return site_; }
    double getMarginalLnL() const { return marginalLnL_; }
    double getMaxIIProb() const { return maxIIProb_; }
    double getMaxProb() const { return maxProb_; }
    static std::string getHeader() {                                         // :119-121
        return std::string("[HEADER]") + Constants::DELIM + "site" + Constants::DELIM + "ln[P(D|theta,alpha)]" +
               Constants::DELIM + "P(A=a,B=b|D,theta,alpha)";
    }
    std::string toString() const {                                           // :124-139
        std::string s = "Result";
        s += Constants::DELIM;
        s += std::to_string(site_ + 1);
        s += Constants::DELIM;
        s += javaDoubleToString(Utils::roundDouble(marginalLnL_, sigDigits_));
        for (size_t i = 0; i < aboveThreshProbs_.size(); ++i) {
            s += Constants::DELIM;
            s += aboveThreshSubs_[i];
            s += Constants::SUB_PROB_DELIM;
            s += javaDoubleToString(Utils::roundDouble(aboveThreshProbs_[i], sigDigits_));
        }
        return s;
    }

private:
    void sortByProbImpl() {                                                  // :96-108, stable descending bubble sort
        auto& p = aboveThreshProbs_;
        auto& sub = aboveThreshSubs_;
        for (size_t i = 0; i + 1 < p.size(); ++i)
            for (size_t j = 1; j < p.size() - i; ++j)
                if (p[j - 1] < p[j]) { std::swap(p[j - 1], p[j]); std::swap(sub[j - 1], sub[j]); }
    }
    int site_;
    double marginalLnL_;
    std::vector<double> aboveThreshProbs_;
    std::vector<std::string> aboveThreshSubs_;
    double maxIIProb_, maxProb_;
    int sigDigits_;
};

// Flattened inputs the Java host already holds after SubRecon.init() (SubRecon.java:157-268).
struct Tree {
    int32_t nNodes = 0, root = 0;
    std::vector<int32_t> childOff, childIdx, leafRow;   // CSR children in Newick order; alignment row per tip, -1 inside
    std::vector<double> branchLength;                   // to the parent
};
struct Alignment {
    int32_t nTaxa = 0;
    int64_t nSites = 0;
    std::vector<uint8_t> states;                        // [taxon][site], 0..19 or SR_STATE_MISSING — or, with `letters`, the
    bool letters = false;                               // residue characters themselves (translated on the device)
};
struct Model {                                          // PAL MatrixExponential.{Eval,Evec,Ievc}
    double Eval[20], Evec[400], Ievc[400];
};

class JointBranchReconstruction {
public:
    static constexpr int MAX_COMPACT_CAP = 64;     // above this many kept entries per column the full matrix is the smaller result
    // JointBranchReconstruction(alignment, tree, model, pi, logPi, rateDist, threshold, sigDigits, sortByProb, sanityCheck, site)
    // — logPi is derived on the device; `site` becomes the (site0, n) range of call(). The call sequence below is the
    // Java host's (java/subrecon/recon/NativeRecon.java + java/SubRecon.run.patch.java), statement for statement:
    // sr_multi_create, "strict_root", set_model, set_rates, set_tree, then per call "input_chars" / "compact_cap",
    // sr_host_alloc'ed buffers and sr_multi_run_streamed.
    JointBranchReconstruction(const Alignment& alignment, const Tree& tree, const Model& model, const double pi[20],
                              const std::vector<double>& rates, double threshold = Constants::DEFAULT_PRINT_THRESHOLD,
                              int sigDigits = Constants::DEFAULT_SIG_DIGITS, bool sortByProb = true,
                              bool sanityCheck = false, int nDevices = 1)
        : threshold_(threshold), sigDigits_(sigDigits), sortByProb_(sortByProb), sanityCheck_(sanityCheck),
          nTaxa_(alignment.nTaxa), nSites_(alignment.nSites), letters_(alignment.letters) {
        if (sr_multi_create(&m_, nDevices, nullptr) != SR_OK) throw std::runtime_error(sr_multi_last_error(nullptr));
        try {
            ck(sr_multi_set_option(m_, "strict_root", 1));
            ck(sr_multi_set_model(m_, model.Eval, model.Evec, model.Ievc, pi));
            ck(sr_multi_set_rates(m_, (int)rates.size(), rates.data()));
            ck(sr_multi_set_tree(m_, tree.nNodes, tree.childOff.data(), tree.childIdx.data(), tree.branchLength.data(),
                                 tree.leafRow.data(), tree.root));
            const size_t bytes = (size_t)alignment.nTaxa * (size_t)alignment.nSites;
            states_ = static_cast<uint8_t*>(sr_host_alloc(bytes));                   // NativeRecon.pinnedLetters
            if (!states_) throw std::runtime_error("sr_host_alloc failed");
            std::memcpy(states_, alignment.states.data(), bytes);
        } catch (...) {
            sr_host_free(states_);
            sr_multi_destroy(m_);
            throw;
        }
    }
    JointBranchReconstruction(const JointBranchReconstruction&) = delete;
    JointBranchReconstruction& operator=(const JointBranchReconstruction&) = delete;
    ~JointBranchReconstruction() { sr_host_free(states_); sr_multi_destroy(m_); }

    // Callable.call() for every column in [site0, site0 + n): results in column order (Future.get(), SubRecon.java:123-126)
    std::vector<SiteResult> call(int64_t site0 = 0, int64_t n = -1) {
        if (n < 0) n = nSites_ - site0;
        ck(sr_multi_set_option(m_, "input_chars", letters_ ? 1 : 0));
        const int cap = sr_compact_cap_for(threshold_);
        const bool useCompact = cap <= MAX_COMPACT_CAP && !sanityCheck_;             // -debug sums all 400 entries
        double* lnl = static_cast<double*>(sr_host_alloc(sizeof(double) * (size_t)n));
        double* joint = nullptr;
        unsigned char* compact = nullptr;
        size_t stride = 0;
        std::vector<SiteResult> out;
        try {
            if (useCompact) {
                ck(sr_multi_set_option(m_, "compact_cap", cap));
                stride = sr_compact_stride(cap);
                compact = static_cast<unsigned char*>(sr_host_alloc(stride * (size_t)n));
            } else {
                joint = static_cast<double*>(sr_host_alloc(sizeof(double) * 400 * (size_t)n));
            }
            if (!lnl || !(useCompact ? (void*)compact : (void*)joint)) throw std::runtime_error("sr_host_alloc failed");
            ck(sr_multi_run_streamed(m_, nTaxa_, nSites_, states_, nSites_, site0, n, lnl, joint,
                                     reinterpret_cast<sr_compact*>(compact), threshold_));
            out.reserve((size_t)n);
            for (int64_t i = 0; i < n; ++i) {
                if (useCompact) {
                    out.push_back(SiteResult::fromCompact((int)(site0 + i), compact + stride * (size_t)i, cap, threshold_, sortByProb_, sigDigits_));
                    continue;
                }
                const double* p = joint + (size_t)i * 400;
                if (sanityCheck_) {                                          // checkSumToOne, JointBranchReconstruction.java:157-166
                    double sum = 0.0;
                    for (int e = 0; e < 400; ++e) sum += p[e];
                    if (sum < 1.0 - Constants::EPSILON || sum > 1.0 + Constants::EPSILON)
                        throw std::runtime_error("ERROR: Failed sanity check. Sum of posterior probs != 1.0. sum=" + javaDoubleToString(sum));
                }
                out.emplace_back((int)(site0 + i), lnl[(size_t)i], p, threshold_, sortByProb_, sigDigits_);
            }
        } catch (...) {
            sr_host_free(lnl); sr_host_free(joint); sr_host_free(compact);
            throw;
        }
        sr_host_free(lnl); sr_host_free(joint); sr_host_free(compact);
        return out;
    }
    sr_multi* contexts() { return m_; }

private:
    void ck(int rc) { if (rc != SR_OK) throw std::runtime_error(sr_multi_last_error(m_)); }
    sr_multi* m_ = nullptr;
    uint8_t* states_ = nullptr;
    double threshold_;
    int sigDigits_;
    bool sortByProb_, sanityCheck_;
    int32_t nTaxa_;
    int64_t nSites_;
    bool letters_;
};

// The ordered join + print loop of SubRecon.run() (SubRecon.java:122-150); returns the stdout lines.
inline std::vector<std::string> runColumns(const std::vector<SiteResult>& results, bool verbose, double threshold,
                                           bool singleSite = false) {
    std::vector<std::string> lines;
    bool printingSites = false;
    double totalLnL = 0.0;
    for (const SiteResult& r : results) {
        totalLnL += r.getMarginalLnL();
        if (verbose || (r.getMaxIIProb() <= 1. - threshold && r.getMaxProb() >= threshold)) {
            printingSites = true;
            lines.push_back(r.toString());
        }
    }
    char buf[128];
    if (!singleSite) { std::snprintf(buf, sizeof(buf), "Total lnL: %.10f", totalLnL); lines.emplace_back(buf); }
    if (!printingSites) {
        std::snprintf(buf, sizeof(buf), "0 sites have non-identical substitution probabilities greater than threshold value (threshold=%.5f)", threshold);
        lines.emplace_back(buf);
        lines.emplace_back("The options -threshold, -nosort and -verbose can be used to control output detail");
    }
    return lines;
}

}  // namespace subrecon
#endif  // SUBRECON_B200_HOST_HPP
