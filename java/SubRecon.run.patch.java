// Replacement for the column loop of SubRecon.run() — src/subrecon/SubRecon.java:102-143 — with libsubrecon_b200.
// Everything around it (init(), the banner, getModelInstance, the "0 sites have ..." message, lines 145-150) is untouched.
// Not compiled here (no JDK in the build image); tests/cpp/host_mirror.cpp performs the same calls in the same order.

        boolean printingSites = false;
        try (NativeRecon nr = new NativeRecon(nThreads)) {                     // -threads N now means "N GPUs" (<= 0: all)
            nr.setOption("strict_root", 1);                                    // a tip as root child with K > 1: fail once, not per column
            nr.setModel(getModelInstance(comArgs.getModelID(), pi), pi);       // same model constructor as line 113
            nr.setRates(rateDist);
            nr.setTree(tree, alignment);                                       // validates taxon names up front
            ByteBuffer states = NativeRecon.pinnedLetters(alignment);
            try {
                int L = alignment.getLength();
                if (site > -1) {                                               // -site n (lines 102-107): one column, no total
                    SiteResult result = nr.reconstruct(states, L, site, 1, threshold, sortByProb, sigDigits)[0];
                    if (verbose || (result.getMaxIIProb() <= 1. - threshold && result.getMaxProb() >= threshold)) {
                        printingSites = true;
                        System.out.println(result);
                    }
                } else {
                    double totalLnL = 0.0;                                     // summed in column order, as line 137 does
                    final int SLICE = 1 << 22;                                 // result buffers of a few hundred MB at a time
                    for (int s0 = 0; s0 < L; s0 += SLICE) {
                        SiteResult[] results = nr.reconstruct(states, L, s0, Math.min(SLICE, L - s0), threshold, sortByProb, sigDigits);
                        for (SiteResult result : results) {                    // unchanged print loop, lines 137-141
                            totalLnL += result.getMarginalLnL();
                            if (verbose || (result.getMaxIIProb() <= 1. - threshold && result.getMaxProb() >= threshold)) {
                                printingSites = true;
                                System.out.println(result);
                            }
                        }
                    }
                    System.out.printf("Total lnL: %.10f%n", totalLnL);
                }
            } finally {
                NativeRecon.hostFree(states);
            }
        } catch (ReflectiveOperationException e) {
            throw new RuntimeException("cannot read PAL's eigensystem (pal.substmodel.MatrixExponential)", e);
        }
