package subrecon.recon;

import java.lang.reflect.Field;
import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.nio.charset.StandardCharsets;
import java.util.ArrayDeque;
import java.util.ArrayList;
import java.util.IdentityHashMap;
import java.util.List;

import pal.alignment.Alignment;
import pal.substmodel.AminoAcidModel;
import pal.substmodel.RateDistribution;
import pal.tree.Node;
import pal.tree.Tree;

/**
 * JNI surface over libsubrecon_b200.so (include/subrecon_b200.h) for SubRecon's column loop
 * (src/subrecon/SubRecon.java:109-142). One instance drives every GPU of the box from the one JVM
 * (sr_multi_*: a context and a host thread per device inside the library; columns are split into
 * contiguous blocks, results land in place).
 *
 * NOT COMPILED IN THE REPOSITORY THAT SHIPS IT (its build image has no JDK). The C side, jni/subrecon_jni.c, is
 * compiled there against a mock of jni.h; the call sequence of {@link #reconstruct} is mirrored statement by statement
 * in tests/cpp/host_mirror.cpp, which runs against the real library on a GPU.
 */
public final class NativeRecon implements AutoCloseable {
    static { System.loadLibrary("subrecon_jni"); }

    /** Above this many kept entries per column the full 20x20 matrix is the smaller result. */
    private static final int MAX_COMPACT_CAP = 64;

    private long handle;                                       // sr_multi*
    private int nTaxa;

    // ---- natives: one per sr_* entry the host needs (jni/subrecon_jni.c)
    private static native long   multiCreate(int nDevices);                                     // sr_multi_create (<= 0: all)
    private static native void   multiDestroy(long h);                                          // sr_multi_destroy
    private static native String multiLastError(long h);                                        // sr_multi_last_error
    private static native int    multiCount(long h);                                            // sr_multi_count
    private static native int    multiSetModel(long h, double[] eval, double[] evec, double[] ievc, double[] pi);
    private static native int    multiSetRates(long h, double[] rates);
    private static native int    multiSetTree(long h, int[] childOff, int[] childIdx, double[] blen, int[] leafRow, int root);
    private static native int    multiSetOption(long h, String name, long value);               // sr_multi_set_option
    private static native int    multiRunStreamed(long h, int nTaxa, long nSites, ByteBuffer states, long pitch, long site0,
                                                  long n, ByteBuffer lnl, ByteBuffer joint, ByteBuffer compact,
                                                  int compactCap, double threshold);            // sr_multi_run_streamed
    private static native int    compactCapFor(double threshold);                               // sr_compact_cap_for
    private static native long   compactStride(int cap);                                        // sr_compact_stride
    public  static native ByteBuffer hostAlloc(long bytes);                                     // sr_host_alloc -> direct buffer
    public  static native void       hostFree(ByteBuffer b);                                    // sr_host_free

    /** @param nDevices GPUs to use; <= 0 = every visible device (what -threads meant for the CPU pool, SubRecon.java:109) */
    public NativeRecon(int nDevices) {
        handle = multiCreate(nDevices);
        if (handle == 0) throw new RuntimeException(multiLastError(0));
    }

    private void ck(int rc) {
        if (rc != 0) {
            String msg = multiLastError(handle);
            throw new RuntimeException(msg == null || msg.isEmpty() ? "libsubrecon_b200: error " + rc + " (bad argument at the JNI boundary)" : msg);
        }
    }

    public int deviceCount() { return multiCount(handle); }
    public void setOption(String name, long value) { ck(multiSetOption(handle, name, value)); }

    /**
     * PAL keeps the eigensystem in private fields of MatrixExponential; they are populated by any setDistance call and
     * are identical on every later call (SURVEY.md Appendix B.2/B.3). `pi` is SubRecon's field (SubRecon.java:222).
     */
    public void setModel(AminoAcidModel model, double[] pi) throws ReflectiveOperationException {
        model.setDistance(0.1);
        Field f = pal.substmodel.AbstractRateMatrix.class.getDeclaredField("matrixExp_");
        f.setAccessible(true);
        Object me = f.get(model);
        ck(multiSetModel(handle, (double[]) field(me, "Eval"), flat((double[][]) field(me, "Evec")), flat((double[][]) field(me, "Ievc")), pi));
    }

    private static Object field(Object o, String name) throws ReflectiveOperationException {
        Field f = o.getClass().getDeclaredField(name);
        f.setAccessible(true);
        return f.get(o);
    }

    private static double[] flat(double[][] m) {
        if (m.length != 20) throw new IllegalArgumentException("expected a 20x20 matrix");
        double[] out = new double[400];
        for (int i = 0; i < 20; i++) System.arraycopy(m[i], 0, out, 20 * i, 20);
        return out;
    }

    /** rateDist.getRate(i), i < getNumberOfRates() (JointBranchReconstruction.java:86-87); class probabilities are 1/K there. */
    public void setRates(RateDistribution rd) {
        double[] r = new double[rd.getNumberOfRates()];
        for (int i = 0; i < r.length; i++) r[i] = rd.getRate(i);
        ck(multiSetRates(handle, r));
    }

    /**
     * CSR flattening of the PAL tree: nodes numbered in pre-order from the root (0), children in Newick order — the order
     * downTreeMarginal visits them (JointBranchReconstruction.java:208-210) — branch length to the parent per node, the
     * alignment row per tip. A taxon of the tree that the alignment does not hold makes the reference fail on every
     * column (whichIdNumber -> -1 -> ArrayIndexOutOfBounds in getData, AdvancedAlignmentAminoAcid.java:35-36); here it is
     * rejected once, up front.
     */
    public void setTree(Tree tree, Alignment aln) {
        List<Node> nodes = new ArrayList<Node>();
        IdentityHashMap<Node, Integer> number = new IdentityHashMap<Node, Integer>();   // (PAL's own node numbers stay untouched)
        ArrayDeque<Node> stack = new ArrayDeque<Node>();
        stack.push(tree.getRoot());
        while (!stack.isEmpty()) {                                   // iterative: a 5,000-taxon caterpillar is 2,500 levels deep
            Node v = stack.pop();
            number.put(v, nodes.size());
            nodes.add(v);
            for (int i = v.getChildCount() - 1; i >= 0; i--) stack.push(v.getChild(i));
        }
        int n = nodes.size();
        int[] childOff = new int[n + 1], childIdx = new int[n - 1], leafRow = new int[n];
        double[] blen = new double[n];
        int e = 0;
        for (int v = 0; v < n; v++) {
            Node node = nodes.get(v);
            childOff[v] = e;
            for (int i = 0; i < node.getChildCount(); i++) childIdx[e++] = number.get(node.getChild(i));
            blen[v] = node.getBranchLength();
            leafRow[v] = -1;
            if (node.isLeaf()) {
                String name = node.getIdentifier().getName();
                int row = aln.whichIdNumber(name);
                if (row < 0) throw new IllegalArgumentException("ERROR: taxon '" + name + "' of the tree is not in the alignment");
                leafRow[v] = row;
            }
        }
        childOff[n] = e;
        nTaxa = aln.getSequenceCount();
        ck(multiSetTree(handle, childOff, childIdx, blen, leafRow, 0));
    }

    /**
     * The alignment as the library wants it for "input_chars": one row of residue LETTERS per sequence in page-locked
     * memory; the device applies pal.datatype.AminoAcids.getState (K0, encode_chars_kernel). Replaces 10^9 calls of
     * getStateBySequenceName at 1,000 taxa x 1M sites (AdvancedAlignmentAminoAcid.java:34-40).
     */
    public static ByteBuffer pinnedLetters(Alignment aln) {
        int n = aln.getSequenceCount(), len = aln.getSiteCount();
        ByteBuffer states = hostAlloc((long) n * len);
        if (states == null) throw new OutOfMemoryError("sr_host_alloc");
        for (int t = 0; t < n; t++) {
            byte[] row = aln.getAlignedSequenceString(t).getBytes(StandardCharsets.ISO_8859_1);
            if (row.length != len) throw new IllegalArgumentException("ragged alignment at sequence " + t);
            for (int s = 0; s < len; s++) states.put((int) ((long) t * len + s), row[s]);   // (absolute put: positions stay 0)
        }
        return states;
    }

    /**
     * The submit/join loop of SubRecon.run() (SubRecon.java:112-126) for columns [site0, site0+n): one SiteResult per
     * column, in column order. At the thresholds people use (default 0.5: at most two entries can pass) the library
     * returns the compact record — exactly what SiteResult keeps of a column (SiteResult.java:55,69-83), 112 bytes instead
     * of 3,208 — and the SiteResult is built through its public constructor from a matrix that holds those entries. For
     * thresholds so low that the record would outgrow the matrix (-threshold 0.0) the 400 probabilities come back.
     */
    public SiteResult[] reconstruct(ByteBuffer states, long nSites, long site0, int n, double threshold, boolean sortByProb, int sigDigits) {
        setOption("input_chars", 1);
        int cap = compactCapFor(threshold);
        boolean useCompact = cap <= MAX_COMPACT_CAP;
        ByteBuffer lnl = hostAlloc(8L * n), joint = null, compact = null;
        long stride = 0;
        if (useCompact) {
            setOption("compact_cap", cap);
            stride = compactStride(cap);
            compact = hostAlloc(stride * n);
        } else {
            joint = hostAlloc(3200L * n);
        }
        try {
            if (lnl == null || (useCompact ? compact : joint) == null) throw new OutOfMemoryError("sr_host_alloc");
            ck(multiRunStreamed(handle, nTaxa, nSites, states, nSites, site0, n, lnl, joint, compact, cap, threshold));
            lnl.order(ByteOrder.nativeOrder());
            SiteResult[] out = new SiteResult[n];
            double[][] probs = new double[20][20];
            if (useCompact) {
                compact.order(ByteOrder.nativeOrder());
                int pairOff = 32, probOff = 32 + (2 * cap + 7) / 8 * 8;
                for (int i = 0; i < n; i++) {
                    int base = (int) (i * stride);                                   // (blocks of < 2^31 bytes; larger runs go in slices)
                    double maxII = compact.getDouble(base + 8), maxP = compact.getDouble(base + 16);
                    int kept = compact.getInt(base + 24);
                    if (kept > cap) throw new IllegalStateException("compact record overflow at site " + (site0 + i + 1));
                    for (double[] row : probs) java.util.Arrays.fill(row, 0.0);
                    for (int k = 0; k < kept; k++) {
                        int ab = compact.getShort(base + pairOff + 2 * k);
                        probs[ab / 20][ab % 20] = compact.getDouble(base + probOff + 8 * k);
                    }
                    // SiteResult recomputes maxIIProb / maxProb from the matrix (SiteResult.java:72-75): give it entries that
                    // carry them when they are below the threshold (and therefore not among the kept ones)
                    if (maxII < threshold) probs[0][0] = Math.max(probs[0][0], maxII);
                    if (maxP < threshold && maxP > maxII) probs[0][1] = Math.max(probs[0][1], maxP);
                    out[i] = new SiteResult((int) (site0 + i), compact.getDouble(base), probs, threshold, sortByProb, sigDigits);
                }
            } else {
                joint.order(ByteOrder.nativeOrder());
                for (int i = 0; i < n; i++) {
                    for (int a = 0; a < 20; a++)
                        for (int b = 0; b < 20; b++) probs[a][b] = joint.getDouble((i * 400 + a * 20 + b) * 8);
                    out[i] = new SiteResult((int) (site0 + i), lnl.getDouble(8 * i), probs, threshold, sortByProb, sigDigits);
                }
            }
            return out;
        } finally {
            if (lnl != null) hostFree(lnl);
            if (joint != null) hostFree(joint);
            if (compact != null) hostFree(compact);
        }
    }

    @Override public void close() {
        if (handle != 0) { multiDestroy(handle); handle = 0; }
    }
}
