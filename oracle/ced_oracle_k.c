/*
 * ced_oracle_k.c -- TEST INFRASTRUCTURE ONLY (same rules as ced_oracle.c: nothing under convolutionalencdec_b200/
 * may link, load or call this file).
 *
 * CPU restatement of the reference's rate-k/n codes with k > 1: ONE shift register of k*K bits that
 * takes k bits per segment (src/convEncode.h:8-18, src/convEncode.c:46-130), a trellis of 2^(k*S) states with 2^k
 * branches into every state (src/viterbiDecoder.c:95-128).
 *
 * Parity status.
 *   PINNED against the unmodified reference compiled with k = 2 parameter headers (oracle/_ref/libced_refk_*.so,
 *   oracle/ref_harness_k.c; tests/test_oracle_k.py): the encoder output, the trellis edge labels, and the path
 *   metrics after EVERY step of the add-compare-select.
 *   UNPINNED: the decoded bytes.  The reference's only k > 1 decoder (generic viterbiDecoderHard) emits garbage at
 *   HEAD -- its register-exchange traceback shifts a uint8_t by (5K-1)k bits (src/viterbiDecoder.h:75-76,
 *   src/viterbiDecoder.c:148).  The traceback restated here is the one the reference's butterfly decoder performs
 *   at `last` (src/viterbiDecoderButterflyk1.c:200-256), whose formulas are written for general k (state >> k,
 *   decision << (S-1)k, byte index t*k/8, k bits per step into the byte) although that file's ACS is k = 1 only:
 *   full traceback from state 0 over the decisions of the pinned ACS.  Round trips (decode(encode(m)) == m, and
 *   under correctable noise) are checked in the tests.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORCK_MAX_STATES 256
#define ORCK_MAX_N 8
#define ORCK_MAX_BRANCH 16

/* src/convEncode.c:163-175: reverse the k*K generator bits so that the LSb taps the bit shifted in last */
static uint32_t orck_reverse(uint64_t gen, int bits)
{
    uint32_t r = 0;
    for (int i = 0; i < bits; i++) {
        r = (r << 1) | (uint32_t)(gen & 1u);
        gen >>= 1;
    }
    return r;
}

/* src/convEncode.c:132-161: generator i contributes bit i */
static uint8_t orck_segment(uint32_t reg, const uint32_t *taps, int n)
{
    uint8_t seg = 0;
    for (int i = 0; i < n; i++)
        seg |= (uint8_t)((__builtin_popcount(reg & taps[i]) & 1) << i);
    return seg;
}

void orck_taps(int K, int k, int n, const uint64_t *g, uint32_t *taps)
{
    for (int i = 0; i < n; i++)
        taps[i] = orck_reverse(g[i], k * K);
}

/*
 * src/convEncode.c:46-130.  Bytes ascending, MSb first; a segment is emitted after every k bits shifted in (:56-97);
 * when `last`, S segments of k zero bits follow and the register returns to 0 (:100-122).  bytesIn * 8 must be a
 * multiple of k when last (:103-106 exits otherwise): returns -1.  `reg` carries the register between calls;
 * sub-segment leftovers (8 % k != 0) are not carried here: every call must hold whole segments.
 */
int orck_encode(int K, int k, int n, const uint64_t *g, uint32_t *reg, const uint8_t *in, int bytesIn, uint8_t *segs,
                int last)
{
    uint32_t taps[ORCK_MAX_N];
    const uint32_t keep = (k * K >= 32) ? 0xFFFFFFFFu : ((1u << (k * K)) - 1u);
    int out = 0, pending = 0;
    if ((bytesIn * 8) % k)
        return -1;
    orck_taps(K, k, n, g, taps);
    for (int b = 0; b < bytesIn; b++)
        for (int bit = 7; bit >= 0; bit--) {
            *reg = ((*reg << 1) | ((in[b] >> bit) & 1u)) & keep;
            if (++pending == k) {
                segs[out++] = orck_segment(*reg, taps, n);
                pending = 0;
            }
        }
    if (last) {
        for (int i = 0; i < K - 1; i++) {
            *reg = (*reg << k) & keep;
            segs[out++] = orck_segment(*reg, taps, n);
        }
        *reg = 0;
    }
    return out;
}

int orck_encode_batch(int K, int k, int n, const uint64_t *g, const uint8_t *in, size_t inStride, int nFrames,
                      int bytesPerFrame, uint8_t *segs, size_t segStride)
{
    for (int f = 0; f < nFrames; f++) {
        uint32_t reg = 0;
        if (orck_encode(K, k, n, g, &reg, in + (size_t)f * inStride, bytesPerFrame, segs + (size_t)f * segStride, 1) < 0)
            return -1;
    }
    return 0;
}

/* viterbiInit, src/viterbiDecoder.c:32-50: edge[e * N + s] = segment on the branch that leaves state s with the k
 * input bits e (shifted in MSb first, src/convEncode.c:19-44) */
void orck_edges(int K, int k, int n, const uint64_t *g, uint8_t *edge)
{
    uint32_t taps[ORCK_MAX_N];
    const int N = 1 << (k * (K - 1)), P = 1 << k;
    orck_taps(K, k, n, g, taps);
    for (int e = 0; e < P; e++)
        for (int s = 0; s < N; s++)
            edge[e * N + s] = orck_segment(((uint32_t)s << k) | (uint32_t)e, taps, n);
}

/*
 * Forward recursion, src/viterbiDecoder.c:95-128, one frame:
 *   destination state d: edgeOut = d % 2^k; the 2^k sources are d / 2^k + edgeIn * 2^((S-1)k)        (:101-111)
 *   path metric = source metric + calcHammingDist(label, rx, n)                                      (:113-115)
 *   the smallest wins, the LOWEST edgeIn on equal metrics (argminPathMetrics: pairwise '<=' trees,  (:118, :287-330)
 *   which keep the left entry)
 *   start metrics 0 / NUM_STATES + 1 (resetViterbiDecoderHard, :236-258); no renormalisation: METRIC_TYPE is sized so
 *   that n * MAX_PKT_LEN_SEGMENTS fits (src/viterbiDecoder.h:52-61)
 * metricsTrace (may be NULL): [T][N] metrics after every step.  surv: [T][N] winning edgeIn per state and step.
 */
static void orck_forward(int K, int k, int n, const uint8_t *edge, const uint8_t *segs, int T, uint8_t *surv,
                         uint32_t *metricsTrace)
{
    const int S = K - 1, N = 1 << (k * S), P = 1 << k, top = (S - 1) * k;
    const uint32_t nmask = (1u << n) - 1u;
    uint32_t cur[ORCK_MAX_STATES], next[ORCK_MAX_STATES];
    cur[0] = 0;
    for (int s = 1; s < N; s++)
        cur[s] = (uint32_t)N + 1u;
    for (int t = 0; t < T; t++) {
        const uint32_t rx = segs[t];
        for (int d = 0; d < N; d++) {
            const int edgeOut = d % P;
            uint32_t best = 0, bestIn = 0;
            for (int edgeIn = 0; edgeIn < P; edgeIn++) {
                const int src = d / P + (edgeIn << top);
                const uint32_t pm = cur[src] + (uint32_t)__builtin_popcount((edge[edgeOut * N + src] ^ rx) & nmask);
                if (edgeIn == 0 || pm < best) { /* strict '<': the lowest edgeIn keeps a tie */
                    best = pm;
                    bestIn = (uint32_t)edgeIn;
                }
            }
            next[d] = best;
            surv[(size_t)t * N + d] = (uint8_t)bestIn;
        }
        memcpy(cur, next, sizeof(uint32_t) * (size_t)N);
        if (metricsTrace)
            memcpy(metricsTrace + (size_t)t * N, cur, sizeof(uint32_t) * (size_t)N);
    }
}

/* path metrics after every step (for pinning against the reference's generic decoder) */
int orck_metrics(int K, int k, int n, const uint64_t *g, const uint8_t *segs, int T, uint32_t *metrics)
{
    const int N = 1 << (k * (K - 1));
    if (N > ORCK_MAX_STATES || (1 << k) > ORCK_MAX_BRANCH)
        return -1;
    uint8_t *edge = (uint8_t *)malloc((size_t)(1 << k) * N), *surv = (uint8_t *)malloc((size_t)T * N);
    orck_edges(K, k, n, g, edge);
    orck_forward(K, k, n, edge, segs, T, surv, metrics);
    free(edge);
    free(surv);
    return 0;
}

/*
 * Whole frames, one-shot: forward recursion above, then the full traceback of
 * src/viterbiDecoderButterflyk1.c:200-256 with its k-generic formulas: start in state 0 (:205), walk the S tail steps
 * without output (:208-223), then for every remaining step t write the k newest state bits into byte t*k/8, filling
 * it from the top (:244-249), and move to the source state (state >> k) | decision << (S-1)k (:252).
 * 8 % k must be 0 and (T - S) * k a multiple of 8.  Output: (T - S) * k / 8 bytes per frame.
 */
int orck_decode_batch(int K, int k, int n, const uint64_t *g, const uint8_t *segs, size_t segStride, int nFrames, int T,
                      uint8_t *out, size_t outStride)
{
    const int S = K - 1, N = 1 << (k * S), P = 1 << k, top = (S - 1) * k;
    if (N > ORCK_MAX_STATES || P > ORCK_MAX_BRANCH || 8 % k || T <= S || ((T - S) * k) % 8)
        return -1;
    uint8_t *edge = (uint8_t *)malloc((size_t)P * N), *surv = (uint8_t *)malloc((size_t)T * N);
    orck_edges(K, k, n, g, edge);
    for (int f = 0; f < nFrames; f++) {
        uint8_t *o = out + (size_t)f * outStride;
        orck_forward(K, k, n, edge, segs + (size_t)f * segStride, T, surv, NULL);
        uint32_t state = 0;
        for (int i = 0; i < S; i++) {
            const int t = T - 1 - i;
            state = (state >> k) | ((uint32_t)surv[(size_t)t * N + state] << top);
        }
        memset(o, 0, (size_t)((T - S) * k / 8));
        for (int i = S; i < T; i++) {
            const int t = T - 1 - i;
            const uint32_t dec = surv[(size_t)t * N + state];
            o[t * k / 8] = (uint8_t)((o[t * k / 8] >> k) | ((state & (uint32_t)(P - 1)) << (8 - k)));
            state = (state >> k) | (dec << top);
        }
    }
    free(edge);
    free(surv);
    return 0;
}
