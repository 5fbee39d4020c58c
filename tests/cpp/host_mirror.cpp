// Test driver for include/subrecon_b200_host.hpp (the C++ mirror of SiteResult / JointBranchReconstruction).
//   host_mirror format            -> formatting cases, no GPU
//   host_mirror run <blob> <threshold> <verbose 0|1> <sd>   -> runs the native path on a flattened problem and prints
//                                    what SubRecon would print (Result lines, Total lnL)
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>

#include "subrecon_b200_host.hpp"

using namespace subrecon;

template <typename T>
static void rd(std::ifstream& f, std::vector<T>& v, size_t n) { v.resize(n); f.read(reinterpret_cast<char*>(v.data()), sizeof(T) * n); }

int main(int argc, char** argv) {
    if (argc >= 2 && !std::strcmp(argv[1], "format")) {
        const double xs[] = {1.0, 0.99, -15.66, 0.0, 1.2e-4, 0.001, 12345678.9, 1e-10, -0.0, 123.0, 0.5, 9999999.0, 1e7, 0.30000000000000004};
        for (double x : xs) std::cout << javaDoubleToString(x) << "\n";
        std::cout << javaDoubleToString(Utils::roundDouble(-2.5, 0)) << " " << javaDoubleToString(Utils::roundDouble(0.125, 2)) << " "
                  << javaDoubleToString(Utils::roundDouble(-15.664999, 2)) << "\n";
        double p[400] = {0};
        p[1 * 20 + 11] = 0.3; p[0] = 0.25; p[2 * 20 + 3] = 0.3; p[5 * 20 + 5] = 0.15;
        SiteResult r(13, -7.004, p, 0.2, true, 2), r2(13, -7.004, p, 0.2, false, 2);
        {   // a compact record (capacity 5) carrying the same column must print the same line
            unsigned char rec[32 + 16 + 40] = {0};
            const double lnl = -7.004, mii = 0.25, mp = 0.3, pr[3] = {0.25, 0.3, 0.3};
            const int32_t kept = 3, inter = 1;
            const int16_t ab[3] = {0, 1 * 20 + 11, 2 * 20 + 3};
            std::memcpy(rec, &lnl, 8); std::memcpy(rec + 8, &mii, 8); std::memcpy(rec + 16, &mp, 8); std::memcpy(rec + 24, &kept, 4); std::memcpy(rec + 28, &inter, 4);
            std::memcpy(rec + 32, ab, 6); std::memcpy(rec + 48, pr, 24);
            SiteResult rc5 = SiteResult::fromCompact(13, rec, 5, 0.2, true, 2);
            if (rc5.toString() != r.toString() || rc5.getMaxIIProb() != r.getMaxIIProb() || rc5.getMaxProb() != r.getMaxProb()) { std::cout << "COMPACT MISMATCH\n"; return 4; }
        }
        std::cout << r.toString() << "\n" << r2.toString() << "\n" << SiteResult::getHeader() << "\n";
        std::cout << javaDoubleToString(r.getMaxIIProb()) << " " << javaDoubleToString(r.getMaxProb()) << "\n";
        for (const auto& l : runColumns({r}, false, 0.5)) std::cout << l << "\n";
        return 0;
    }
    if (argc >= 6 && !std::strcmp(argv[1], "run")) {
        std::ifstream f(argv[2], std::ios::binary);
        int64_t hdr[8];   // n_nodes, n_edges, n_taxa, n_sites, K, root, alignment holds letters (0/1), sanity check (0/1)
        f.read(reinterpret_cast<char*>(hdr), sizeof(hdr));
        Tree t; Alignment a; Model m; std::vector<double> pi, rates, tmp;
        t.nNodes = (int32_t)hdr[0]; t.root = (int32_t)hdr[5];
        rd(f, t.childOff, hdr[0] + 1); rd(f, t.childIdx, hdr[1]); rd(f, t.branchLength, hdr[0]); rd(f, t.leafRow, hdr[0]);
        a.nTaxa = (int32_t)hdr[2]; a.nSites = hdr[3]; a.letters = hdr[6] != 0;
        rd(f, a.states, (size_t)hdr[2] * hdr[3]);
        rd(f, tmp, 840);
        std::memcpy(m.Eval, tmp.data(), 160); std::memcpy(m.Evec, tmp.data() + 20, 3200); std::memcpy(m.Ievc, tmp.data() + 420, 3200);
        pi.assign(tmp.begin() + 820, tmp.end());
        rd(f, rates, hdr[4]);
        const double threshold = std::atof(argv[3]);
        const bool verbose = std::atoi(argv[4]) != 0;
        const int sd = std::atoi(argv[5]);
        try {
            JointBranchReconstruction jbr(a, t, m, pi.data(), rates, threshold, sd, true, /*sanityCheck=*/hdr[7] != 0);
            std::cout << SiteResult::getHeader() << "\n";
            for (const auto& l : runColumns(jbr.call(), verbose, threshold)) std::cout << l << "\n";
        } catch (const std::exception& e) {
            std::cout << "EXCEPTION: " << e.what() << "\n";
            return 3;
        }
        return 0;
    }
    std::fprintf(stderr, "usage: host_mirror format | run <blob> <threshold> <verbose> <sd>\n");
    return 2;
}
