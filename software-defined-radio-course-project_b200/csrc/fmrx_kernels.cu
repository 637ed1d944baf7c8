// fmrx_kernels.cu -- hand-written sm_100a kernels of the FM receive chain.
//
//   K1 k_rf_demod      u8 IQ unpack + RF low-pass/decimate (I,Q) + FM discriminator
//                      (reference: src/iofunc.cpp:62-69, src/project.cpp:57-69,
//                       src/filter.cpp:84-92 with up=1, src/filter.cpp:110-132)
//   K2 k_bandpass_pair pilot and stereo-band band-pass FIRs (src/project.cpp:162,165)
//   K3 k_pll           PLL recurrence, one warp per capture (src/filter.cpp:157-171)
//   K4 k_audio         NCO cosine, mixer, mono+stereo polyphase low-pass with the
//                      shared-state block quirk, mono delay, L/R combine, s16 pack
//                      (src/filter.cpp:170,176-199, src/project.cpp:146-193)
//
// Data layout: every per-capture stream is contiguous ([capture][sample]); IF-rate
// arrays carry `if_off` samples of history in front of the chunk so that FIR taps
// that reach before the chunk start are plain negative indices.
//
// FIR arithmetic is acc = fadd(acc, fmul(c[k], x)) from +0 with k ascending --
// the reference's exact sequence -- so one MAC costs an FMUL and an FADD: the
// attainable ceiling of these kernels is half the FFMA peak by construction.
#include "fmrx_internal.h"

namespace fmrx {

// ============================================================================
// K1: u8 unpack + RF FIR (decimating) + FM discriminator
// ============================================================================
//
// A tile computes RF_COMPUTED consecutive IF outputs; the first one is the
// discriminator's "previous sample" halo, so RF_COMPUTED-1 demod samples are
// produced.  The u8 IQ window of the tile is converted to float ONCE while it
// is staged into shared memory, de-interleaved into an I plane and a Q plane,
// each stored phase-major (index m -> row m%decim, column m/decim): output o
// reads x[o*decim + e] = row e%decim, column o + e/decim, so the 32 lanes of a
// warp (consecutive o) hit consecutive banks for any decimation factor.

constexpr int RF_THREADS = 256;
constexpr int RF_R = 2;                                 // outputs per thread
constexpr int RF_COMPUTED = RF_THREADS * RF_R;          // 512
constexpr int RF_REAL = RF_COMPUTED - 1;                // 511 demod samples per tile

static inline int rf_row_stride(int T, int decim)
{
    const int cols = RF_COMPUTED + (T + decim - 1) / decim + 1;
    const int want = (32 + decim - 1) / decim;          // spreads the staging stores over banks
    int rs = cols;
    while ((rs & 31) != (want & 31))
        rs++;
    return rs;
}

static inline size_t rf_smem_bytes(int T, int decim)
{
    return sizeof(float) * ((size_t)2 * decim * rf_row_stride(T, decim) + 2 * RF_COMPUTED + T);
}

__global__ void __launch_bounds__(RF_THREADS) k_rf_demod(const RfDemodArgs a, const int rs)
{
    extern __shared__ float smem[];
    const int T = a.T, d = a.decim;
    float *s_i = smem;                       // [d][rs]
    float *s_q = s_i + d * rs;               // [d][rs]
    float *o_i = s_q + d * rs;               // [RF_COMPUTED]
    float *o_q = o_i + RF_COMPUTED;
    float *s_c = o_q + RF_COMPUTED;          // [T]

    const int c = blockIdx.y;
    const int tid = threadIdx.x;
    const int n0 = blockIdx.x * RF_REAL;     // first demod sample of the tile
    const long long n_pairs = (long long)a.n_if * d;
    // chunk-local pair index of staged element l:  m = m_base + l
    const long long m_base = (long long)(n0 - 1) * d - (T - 1);
    const int W = (RF_COMPUTED - 1) * d + T;

    const uint8_t *iq = a.iq + (size_t)c * a.iq_stride;
    const uint8_t *hist = a.hist + (size_t)c * 2 * a.hist_pairs;

    for (int k = tid; k < T; k += RF_THREADS)
        s_c[k] = a.taps[k];

    for (int l = tid; l < W; l += RF_THREADS) {
        const long long m = m_base + l;
        uint32_t v = 0x8080u;                // (128,128) -> 0.0f, 0.0f
        if (m >= 0) {
            if (m < n_pairs)
                v = *reinterpret_cast<const uint16_t *>(iq + 2 * m);
        } else {
            const long long h = a.hist_pairs + m;
            if (h >= 0)
                v = *reinterpret_cast<const uint16_t *>(hist + 2 * h);
        }
        const int row = l % d, col = l / d;
        s_i[row * rs + col] = unpack_u8(v & 0xffu);
        s_q[row * rs + col] = unpack_u8(v >> 8);
    }
    __syncthreads();

    float acc_i[RF_R], acc_q[RF_R];
#pragma unroll
    for (int r = 0; r < RF_R; r++) {
        acc_i[r] = 0.0f;
        acc_q[r] = 0.0f;
    }
    // tap k multiplies x[o*d + (T-1-k)]
    int e = T - 1;
    int row = e % d, cb = e / d;
    for (int k = 0; k < T; k++) {
        const float ck = s_c[k];
        const int idx = row * rs + cb + tid;
#pragma unroll
        for (int r = 0; r < RF_R; r++) {
            acc_i[r] = fadd(acc_i[r], fmul(ck, s_i[idx + r * RF_THREADS]));
            acc_q[r] = fadd(acc_q[r], fmul(ck, s_q[idx + r * RF_THREADS]));
        }
        if (--row < 0) {
            row = d - 1;
            cb--;
        }
    }
#pragma unroll
    for (int r = 0; r < RF_R; r++) {
        o_i[tid + r * RF_THREADS] = acc_i[r];
        o_q[tid + r * RF_THREADS] = acc_q[r];
    }
    __syncthreads();

    float *demod = a.demod + (size_t)c * a.if_stride + a.if_off;
#pragma unroll
    for (int r = 0; r < RF_R; r++) {
        const int o = tid + r * RF_THREADS;
        const int n = n0 - 1 + o;
        if (o >= 1 && n < a.n_if) {
            demod[n] = fm_discriminate(o_i[o], o_q[o], o_i[o - 1], o_q[o - 1]);
            if (a.i_ds) {
                const size_t g = (size_t)c * a.stage_stride + a.stage_off + n;
                a.i_ds[g] = o_i[o];
                a.q_ds[g] = o_q[o];
            }
        }
    }
}

cudaError_t launch_rf_demod(const RfDemodArgs &a, int n_captures, cudaStream_t s)
{
    static size_t configured = 0;
    const size_t smem = rf_smem_bytes(a.T, a.decim);
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(k_rf_demod, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess)
            return e;
        configured = smem;
    }
    const int tiles = (a.n_if + RF_REAL - 1) / RF_REAL;
    dim3 grid(tiles, n_captures);
    k_rf_demod<<<grid, RF_THREADS, smem, s>>>(a, rf_row_stride(a.T, a.decim));
    return cudaGetLastError();
}

// ============================================================================
// K2: pilot + stereo-band band-pass pair (no decimation)
// ============================================================================
// One shared demod tile feeds both filters: per staged sample 2 MACs.

constexpr int BP_THREADS = 256;
constexpr int BP_R = 4;
constexpr int BP_TILE = BP_THREADS * BP_R;              // 1024 outputs per tile

__global__ void __launch_bounds__(BP_THREADS) k_bandpass_pair(const BandpassArgs a)
{
    extern __shared__ float smem[];
    const int T = a.T;
    float *s_x = smem;                       // [BP_TILE + T - 1]
    float *s_p = s_x + BP_TILE + T - 1;      // [T] pilot taps
    float *s_c = s_p + T;                    // [T] channel taps

    const int c = blockIdx.y, tid = threadIdx.x;
    const int n0 = blockIdx.x * BP_TILE;
    const float *demod = a.demod + (size_t)c * a.if_stride + a.if_off;

    for (int k = tid; k < T; k += BP_THREADS) {
        s_p[k] = a.taps_pilot[k];
        s_c[k] = a.taps_chan[k];
    }
    for (int l = tid; l < BP_TILE + T - 1; l += BP_THREADS) {
        const int n = n0 - (T - 1) + l;      // >= -(T-1): history in front of if_off
        s_x[l] = (n < a.n_if) ? demod[n] : 0.0f;
    }
    __syncthreads();

    float ap[BP_R], ac[BP_R];
#pragma unroll
    for (int r = 0; r < BP_R; r++) {
        ap[r] = 0.0f;
        ac[r] = 0.0f;
    }
    for (int k = 0; k < T; k++) {
        const float cp = s_p[k], cc = s_c[k];
        const int idx = tid + (T - 1 - k);
#pragma unroll
        for (int r = 0; r < BP_R; r++) {
            const float x = s_x[idx + r * BP_THREADS];
            ap[r] = fadd(ap[r], fmul(cp, x));
            ac[r] = fadd(ac[r], fmul(cc, x));
        }
    }
    float *pilot = a.pilot + (size_t)c * a.pilot_stride;
    float *chan = a.chan + (size_t)c * a.if_stride + a.if_off;
#pragma unroll
    for (int r = 0; r < BP_R; r++) {
        const int n = n0 + tid + r * BP_THREADS;
        if (n < a.n_if) {
            pilot[n] = ap[r];
            chan[n] = ac[r];
        }
    }
}

cudaError_t launch_bandpass_pair(const BandpassArgs &a, int n_captures, cudaStream_t s)
{
    const size_t smem = sizeof(float) * ((size_t)BP_TILE + 3 * a.T);
    dim3 grid((a.n_if + BP_TILE - 1) / BP_TILE, n_captures);
    k_bandpass_pair<<<grid, BP_THREADS, smem, s>>>(a);
    return cudaGetLastError();
}

// ============================================================================
// K3: PLL recurrence (src/filter.cpp:157-171)
// ============================================================================
// The recurrence is one dependent chain per capture, so its latency bounds the
// throughput of the whole receive chain; fmrx_pll_core.h holds the low-latency
// formulation of one step and the measurements behind it.  One warp per capture.
//
// Lanes as value speculation.  The longest piece of a step is everything that hangs
// off the new trigArg: its sin/cos (Cody-Waite + two polynomials), their float
// roundings, the wrapped angle.  But trigArg = fl32(w*trigOffset + phaseEst) lives on
// the float grid of its binade (spacing 2^-7 .. 0.5 rad after the first second), and
// one loop-filter update moves phaseEst by far less than that: the new trigArg is one
// of a handful of grid points around fl32(w*trigOffset + phaseEst_previous).  So as
// soon as phaseEst of step t-1 is known, the 32 lanes evaluate make_feedback() for the
// 32 grid points G_c-16 .. G_c+15 -- SIMT, the same instructions a single evaluation
// costs -- while the chain proceeds through the atan2 shortcut and the loop filter of
// step t.  When step t has its s = w*trigOffset + phaseEst, the grid index
// G = rint(s/ulp) picks the lane (one DFMA, one IADD, a SHFL per word): the selected
// values ARE make_feedback(trigArg), bit for bit.  The dependent chain per step drops
// from ~250 to ~180 cycles: select, float products, FMA residuals, 5 DP operations,
// conversion, loop filter, conversion.
//
// Everything that does not depend on the recurrence is produced 32 samples at a time,
// one lane per sample -- the coalesced pilot load, (double)x, the IEEE reciprocal 1/x,
// the half-turn flag and w*trigOffset -- and parked in shared memory; trigArg of each
// step is parked there by lane 0 and written out coalesced per group.  Steps run
// speculatively in groups of 32 from a register checkpoint: guards (x normal,
// roundings tiny, angle clear of the +-pi seam, binade unchanged, grid point among the
// 32 candidates) only accumulate into a flag, and a group with a failed guard is redone
// step by step with the checked/generic step.  Only trigArg leaves the chain; the NCO
// output cos(trigArg*scale+adjust) is evaluated in K4.

struct __align__(16) PllSlotA {      // per sample u
    float x;                         // pilot sample u
    int turn_hi_next;                // high word of 2.0/0.0 for sample u+1 (x < 0: half a turn)
    double xd;                       // (double)x of sample u
};
struct __align__(16) PllSlotB {
    double v;                        // w * trigOffset after step u (:166-167)
    double inv_x_next;               // 1/(double)x of sample u+1, IEEE divide
};

__device__ __forceinline__ double shfl_d(double v, int src)
{
    return __hiloint2double(__shfl_sync(0xffffffffu, __double2hiint(v), src),
                            __shfl_sync(0xffffffffu, __double2loint(v), src));
}

__global__ void __launch_bounds__(32) k_pll(const PllArgs a)
{
    using namespace pllcore;
    __shared__ PllSlotA s_a[64];     // ring of two groups, index = sample & 63
    __shared__ PllSlotB s_b[64];
    __shared__ double s_first[2][2]; // per group: {turn, inv_x} of its first sample
    __shared__ double s_ta[32];
    __shared__ float4 s_cand[2][32][2];   // per candidate lane: {cf, sf, cr}, {sr, phi}

    const int c = blockIdx.x;
    const int lane = threadIdx.x;
    const float *p = a.pilot + (size_t)c * a.pilot_stride;
    float *tr = a.trig + (size_t)c * a.if_stride + a.if_off;
    float *st = a.state + 8 * (size_t)c;

    Consts k;
    k.kp = a.prm.kp;
    k.ki = a.prm.ki;
    k.w = a.prm.w;
    const TrigK &K = a.kconst;       // kernel-parameter constant bank: direct DFMA operands

    Chain ch;
    ch.integ = st[0];
    ch.ph = st[1];
    ch.fi = st[2];
    ch.fq = st[3];
    ch.toff = st[5];
    chain_load(ch, k);
    const int n = a.n_if;

    // trigOffset after j steps is min(t0 + j, 2^24) when it starts integer-valued
    const bool regular = toff_is_regular(ch.toff);
    const int t0 = regular ? (int)ch.toff : 0;

    // one lane per sample: off-chain inputs of sample base+lane into the ring
    auto prepare = [&](int base, float pvv) {
        const int u = (base + lane) & 63;
        const double xd = (double)pvv;
        const double inv = 1.0 / xd;                                  // IEEE divide
        const int turn_hi = (pvv < 0.0f) ? 0x40000000 : 0;
        const float toff = (float)min(t0 + base + lane + 1, 16777216);    // exact: <= 2^24
        s_a[u].x = pvv;
        s_a[u].xd = xd;
        s_b[u].v = __dmul_rn(k.w, (double)toff);
        // turn / reciprocal are consumed one sample early (with the feedback prepared for it)
        if (lane > 0) {
            s_a[(u - 1) & 63].turn_hi_next = turn_hi;
            s_b[(u - 1) & 63].inv_x_next = inv;
        } else {
            s_first[(base >> 5) & 1][0] = __hiloint2double(turn_hi, 0);
            s_first[(base >> 5) & 1][1] = inv;
            if (base > 0) {
                s_a[(u - 1) & 63].turn_hi_next = turn_hi;
                s_b[(u - 1) & 63].inv_x_next = inv;
            }
        }
    };

    float pvn = (lane < n) ? p[lane] : 1.0f;
    prepare(0, pvn);
    pvn = (32 + lane < n) ? p[32 + lane] : 1.0f;
    bool stale = false;              // ch's sincos leftovers lag behind ch.tad (after speculative groups)
    int n_groups = 0, n_redone = 0;  // diagnostics: state[6], state[7]
    const double lane_off = (double)(lane - 16);

    for (int base = 0; base < n; base += 32) {
        const int nn = base + 64 + lane;
        const float pvnn = (nn < n) ? p[nn] : 1.0f;
        prepare(base + 32, pvn);
        __syncwarp();
        const int cnt = min(32, n - base);
        const Chain ck = ch;
        bool good = regular && ch.binade != FMRX_DISARMED;
        if (good) {
            const double ulp = ch.ulp, inv_ulp = ch.inv_ulp;
            const unsigned binade = ch.binade;
            float integ = ch.integ, ph = ch.ph;
            const int g = (base >> 5) & 1;
            // The loop is rotated so that the two activities of an iteration depend only on
            // the previous iteration and can be interleaved by the scheduler:
            //   A: pick the candidate that is trigArg(t-1) -> feedback of sample t -> atan2
            //      shortcut and loop filter of sample t -> s(t), phaseEst(t)
            //   B: from phaseEst(t-1), the 32 candidates for trigArg(t), combined with the
            //      inputs of sample t+1
            // "candidates" for trigArg(base-1), which is known: every lane holds the real one
            // "candidates" for trigArg(base-1), which is known: every lane holds the real one
            {
                const Feedback f0 = make_feedback(K, ch.tad, s_first[g][0], s_first[g][1], nullptr, nullptr);
                s_cand[1][lane][0] = make_float4(f0.cf, f0.sf, __int_as_float(__double2loint(f0.cr)),
                                                 __int_as_float(__double2hiint(f0.cr)));
                s_cand[1][lane][1] = make_float4(__int_as_float(__double2loint(f0.sr)), __int_as_float(__double2hiint(f0.sr)),
                                                 __int_as_float(__double2loint(f0.phi)), __int_as_float(__double2hiint(f0.phi)));
            }
            double q = FMRX_RINT_MAGIC;          // grid_index(q) - gc + 16 == 16
            int gc = 0;
            double phd = (double)ph;
            double inv_x = s_first[g][1];        // 1/x of the sample about to run
            for (int t = 0; t < cnt; t++) {
                const PllSlotA sa = s_a[(base + t) & 63];
                const PllSlotB sb = s_b[(base + t) & 63];
                __syncwarp();                    // candidate table of the previous iteration is complete
                // ---- A: the candidate that IS trigArg(t-1): two broadcast LDS.128 ----
                const int idx = grid_index(q) - gc + 16;
                const int src = idx & 31;
                const float4 c0 = s_cand[(t + 1) & 1][src][0];
                const float4 c1 = s_cand[(t + 1) & 1][src][1];
                Feedback fb;
                fb.cf = c0.x;
                fb.sf = c0.y;
                fb.cr = __hiloint2double(__float_as_int(c0.w), __float_as_int(c0.z));
                fb.sr = __hiloint2double(__float_as_int(c1.y), __float_as_int(c1.x));
                fb.phi = __hiloint2double(__float_as_int(c1.w), __float_as_int(c1.z));
                fb.csx = p_mul(fb.cr, inv_x);
                fb.snx = p_mul(fb.sr, inv_x);
                // ---- B (uses phaseEst of the previous step only): candidates for trigArg(t),
                //      with the wrapped angle already turned for sample t+1 ----
                const double qc = grid_round(p_add(sb.v, phd), inv_ulp);
                gc = grid_index(qc);
                {
                    const Feedback f = make_feedback(K, grid_value(p_add(qc, lane_off), ulp),
                                                     __hiloint2double(sa.turn_hi_next, 0), 1.0, nullptr, nullptr);
                    s_cand[t & 1][lane][0] = make_float4(f.cf, f.sf, __int_as_float(__double2loint(f.cr)),
                                                         __int_as_float(__double2hiint(f.cr)));
                    s_cand[t & 1][lane][1] = make_float4(__int_as_float(__double2loint(f.sr)), __int_as_float(__double2hiint(f.sr)),
                                                         __int_as_float(__double2loint(f.phi)), __int_as_float(__double2hiint(f.phi)));
                }
                // ---- A continued ----
                bool ok = (unsigned)idx < 32u;
                const double s = step_front(k, fb, sa.x, sa.xd, sb.v, integ, ph, phd, ok);
                q = grid_round(s, inv_ulp);
                good &= ok && in_binade(s, binade);
                if (lane == 0)
                    s_ta[t] = grid_value(q, ulp);
                inv_x = sb.inv_x_next;
            }
            // the last trigArg must be one of its candidates too (checked like the others)
            good &= (unsigned)(grid_index(q) - gc + 16) < 32u;
            if (good) {
                ch.integ = integ;
                ch.ph = ph;
                ch.toff = (float)min(t0 + base + cnt, 16777216);
                ch.tad = grid_value(q, ulp);
                stale = true;
            }
        }
        n_groups++;
        if (!good) {
            n_redone++;
            ch = ck;
            if (stale) {             // bring the sincos leftovers up to date with ch.tad
                chain_refresh(ch);
                stale = false;
            }
            for (int t = 0; t < cnt; t++) {
                const float ta = chain_step(ch, k, K, s_a[(base + t) & 63].x, nullptr);
                if (lane == 0)
                    s_ta[t] = (double)ta;
            }
        }
        __syncwarp();
        if (lane < cnt)
            tr[base + lane] = __double2float_rn(s_ta[lane]);
        __syncwarp();
        pvn = pvnn;
    }
    if (lane == 0) {
        if (stale)
            chain_refresh(ch);
        float fi, fq;
        chain_feedback(ch, fi, fq);
        st[0] = ch.integ;
        st[1] = ch.ph;
        st[2] = fi;
        st[3] = fq;
        st[5] = ch.toff;
        st[6] = (float)n_groups;     // diagnostics of the last launch
        st[7] = (float)n_redone;
        if (n > 0)
            st[4] = nco_from_trig(__double2float_rn(ch.tad), a.prm.scale, a.prm.adjust);   // :173
    }
}

cudaError_t launch_pll(const PllArgs &a_in, int n_captures, cudaStream_t s)
{
    PllArgs a = a_in;
    a.kconst = pllcore::trig_constants();
    k_pll<<<n_captures, 32, 0, s>>>(a);
    return cudaGetLastError();
}

// ============================================================================
// K4: NCO + mixer + mono/stereo polyphase low-pass + delay + combine + pack
// ============================================================================
// One CTA per tile of kAudioTile audio frames of one block of one capture.
// With r = IF index relative to the block start, output n of the block uses
// taps t=0..T-1 at r = floor(n*D/U) - t with coefficient h[(n*D)%U + t*U]
// (src/filter.cpp:85-87).  The reference keeps ONE state vector for the mono
// and the stereo low-pass (src/project.cpp:114,146,172), so for r < 0
//   mono   reads the PREVIOUS block's mixer output (continuous index r), and
//   stereo reads THIS block's demod tail, demod[B + r].
// The 5-frame mono delay (src/project.cpp:153-159) is mono(n-5) on the
// continuous frame index; the 5 frames before a tile are recomputed here.

__global__ void __launch_bounds__(kAudioTile) k_audio(const AudioArgs a, const int span)
{
    extern __shared__ float smem[];
    const int T = a.T, U = a.U, D = a.D, B = a.if_per_block, NA = a.audio_per_block;
    float *s_dem = smem;                     // [span]
    float *s_mix = s_dem + span;             // [span]
    float *s_tail = s_mix + span;            // [T]   demod[B-(T-1) .. B)
    float *s_mono = s_tail + T;              // [kAudioTile + kMonoDelay]

    const int c = blockIdx.y, tid = threadIdx.x;
    const int tiles_per_block = NA / kAudioTile;
    const int b_local = blockIdx.x / tiles_per_block;
    const int n_lo = (blockIdx.x % tiles_per_block) * kAudioTile;

    const size_t cap = (size_t)c * a.if_stride;
    const long long blk0 = (long long)a.if_off + (long long)b_local * B;   // array index of r = 0
    const float *demod = a.demod + cap;
    const float *chan = a.chan + cap;
    const float *trig = a.trig + cap;

    const int q_first = (n_lo >= kMonoDelay) ? ((n_lo - kMonoDelay) * D) / U
                                             : ((NA - kMonoDelay) * D) / U - B;
    const int r_min = q_first - (T - 1);
    // the last tile of a block also stages the block's final samples (no audio frame of
    // this block uses them, but the stage taps cover the whole block)
    const int r_max = (n_lo + kAudioTile >= NA) ? B - 1 : ((n_lo + kAudioTile - 1) * D) / U;
    // IF samples this tile "owns" for the optional nco/mixer stage taps: those past the
    // previous tile's last staged sample
    const int own_lo = n_lo ? ((n_lo - 1) * D) / U + 1 : 0;
    const int own_hi = r_max + 1;
    const int q_lo = (n_lo * D) / U;          // newest sample of the tile's first frame
    const int count = r_max - r_min + 1;

    for (int i = tid; i < count; i += kAudioTile) {
        const int r = r_min + i;
        const long long g = blk0 + r;
        const float ch = chan[g];
        const float nco = nco_from_trig(trig[g], a.scale, a.adjust);     // src/filter.cpp:170
        const float mx = mix2(ch, nco);                                   // src/filter.cpp:182
        s_dem[i] = demod[g];
        s_mix[i] = mx;
        if (a.nco && r >= own_lo && r < own_hi) {
            const size_t sg = (size_t)c * a.if_stage_stride + a.if_stage_off + (size_t)b_local * B + r;
            a.nco[sg] = nco;
            a.mixer[sg] = mx;
        }
    }
    const bool need_tail = (q_lo - (T - 1)) < 0;
    if (need_tail)
        for (int i = tid; i < T - 1; i += kAudioTile)
            s_tail[i] = demod[blk0 + B - (T - 1) + i];
    __syncthreads();

    // this thread's frame
    const int n = n_lo + tid;
    float am = 0.0f, as = 0.0f;
    {
        const int nd = n * D;
        const int q = nd / U;
        const float *cf = a.coef_pm + (size_t)(nd % U) * T;
        for (int t = 0; t < T; t++) {
            const int r = q - t;
            const float cv = cf[t];
            const float xd = s_dem[r - r_min];
            const float xm = s_mix[r - r_min];
            float x_mono, x_st;
            if (r >= 0) {
                x_mono = xd;
                x_st = xm;
            } else {
                x_mono = xm;                       // previous block's mixer tail
                x_st = s_tail[T - 1 + r];          // this block's demod tail
            }
            am = fadd(am, fmul(cv, x_mono));
            as = fadd(as, fmul(cv, x_st));
        }
    }
    s_mono[tid + kMonoDelay] = am;
    if (tid < kMonoDelay) {
        // frames n_lo-5 .. n_lo-1: always plain demod taps (previous block when n_lo == 0)
        int nd, q;
        if (n_lo >= kMonoDelay) {
            nd = (n_lo - kMonoDelay + tid) * D;
            q = nd / U;
        } else {
            nd = (NA - kMonoDelay + tid) * D;
            q = nd / U - B;
        }
        const float *cf = a.coef_pm + (size_t)(nd % U) * T;
        float ae = 0.0f;
        for (int t = 0; t < T; t++)
            ae = fadd(ae, fmul(cf[t], s_dem[q - t - r_min]));
        s_mono[tid] = ae;
    }
    __syncthreads();

    const float ms = s_mono[tid];                              // mono(n - 5)
    const float left = fmul(fadd(ms, as), 0.5f);               // src/filter.cpp:196
    const float right = fmul(fsub(ms, as), 0.5f);              // src/filter.cpp:197
    const size_t frame = (size_t)b_local * NA + n;
    uint32_t *pcm32 = reinterpret_cast<uint32_t *>(a.pcm + (size_t)c * a.pcm_stride);
    pcm32[frame] = pcm_s16(right) | (pcm_s16(left) << 16);     // R first (src/project.cpp:183-191)

    if (a.mono) {
        const size_t sg = (size_t)c * a.au_stage_stride + a.au_stage_off + frame;
        a.mono[sg] = am;
        a.mono_shift[sg] = ms;
        a.stereo[sg] = as;
        a.left[sg] = left;
        a.right[sg] = right;
    }
}

static inline int audio_span(int T, int U, int D)
{
    return (int)(((long long)(kAudioTile + kMonoDelay) * D + U - 1) / U) + T + 2;
}

int audio_smem_bytes(int T, int U, int D)
{
    return (int)sizeof(float) * (2 * audio_span(T, U, D) + T + kAudioTile + kMonoDelay);
}

cudaError_t launch_audio(const AudioArgs &a, int n_captures, cudaStream_t s)
{
    const int tiles = a.n_blocks * (a.audio_per_block / kAudioTile);
    dim3 grid(tiles, n_captures);
    k_audio<<<grid, kAudioTile, audio_smem_bytes(a.T, a.U, a.D), s>>>(a, audio_span(a.T, a.U, a.D));
    return cudaGetLastError();
}

// ============================================================================
// Operator kernels (one reference operator each; device pointers)
// ============================================================================

__global__ void k_u8_to_f32(const uint8_t *raw, size_t n, float *out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        out[i] = unpack_u8(raw[i]);
}

cudaError_t launch_u8_to_f32(const uint8_t *raw, size_t n, float *out, cudaStream_t s)
{
    if (n == 0)
        return cudaSuccess;
    k_u8_to_f32<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(raw, n, out);
    return cudaGetLastError();
}

// src/filter.cpp:84-92, one thread per kept output.
__global__ void k_resample(float *out, int n_out, const float *state, int state_len,
                           const float *in, const float *coeff, int taps, int up, int down)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_out)
        return;
    const long long nd = (long long)n * down;
    float acc = 0.0f;
    for (int k = (int)(nd % up); k < taps; k += up) {
        const long long j = (nd - k) / up;
        float x;
        if (j >= 0) {
            x = in[j];
        } else {
            const long long sj = state_len + j;
            x = (sj >= 0) ? state[sj] : 0.0f;
        }
        acc = fadd(acc, fmul(coeff[k], x));
    }
    out[n] = acc;
}

cudaError_t launch_resample(float *out, int n_out, const float *state, int state_len,
                            const float *in, int n_in, const float *coeff, int taps,
                            int up, int down, cudaStream_t s)
{
    (void)n_in;
    if (n_out == 0)
        return cudaSuccess;
    k_resample<<<(n_out + 127) / 128, 128, 0, s>>>(out, n_out, state, state_len, in, coeff, taps,
                                                   up, down);
    return cudaGetLastError();
}

__global__ void k_fmdemod(float *out, const float *i_ds, const float *q_ds, int n, float prev_i,
                          float prev_q)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n)
        return;
    const float pi_ = k ? i_ds[k - 1] : prev_i;
    const float pq_ = k ? q_ds[k - 1] : prev_q;
    out[k] = fm_discriminate(i_ds[k], q_ds[k], pi_, pq_);
}

cudaError_t launch_fmdemod(float *out, const float *i_ds, const float *q_ds, int n, float prev_i,
                           float prev_q, cudaStream_t s)
{
    if (n == 0)
        return cudaSuccess;
    k_fmdemod<<<(n + 255) / 256, 256, 0, s>>>(out, i_ds, q_ds, n, prev_i, prev_q);
    return cudaGetLastError();
}

__global__ void k_nco(float *out, const float *trig, size_t n, float scale, float adjust)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        out[i] = nco_from_trig(trig[i], scale, adjust);
}

cudaError_t launch_nco(float *out, const float *trig, size_t n, float scale, float adjust,
                       cudaStream_t s)
{
    if (n == 0)
        return cudaSuccess;
    k_nco<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(out, trig, n, scale, adjust);
    return cudaGetLastError();
}

__global__ void k_mixer(float *out, const float *x, const float *y, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        out[i] = mix2(x[i], y[i]);
}

cudaError_t launch_mixer(float *out, const float *x, const float *y, size_t n, cudaStream_t s)
{
    if (n == 0)
        return cudaSuccess;
    k_mixer<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(out, x, y, n);
    return cudaGetLastError();
}

__global__ void k_lr_extract(float *left, float *right, const float *mono, const float *stereo,
                             size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float m = mono[i], s = stereo[i];
        left[i] = fmul(fadd(m, s), 0.5f);
        right[i] = fmul(fsub(m, s), 0.5f);
    }
}

cudaError_t launch_lr_extract(float *left, float *right, const float *mono, const float *stereo,
                              size_t n, cudaStream_t s)
{
    if (n == 0)
        return cudaSuccess;
    k_lr_extract<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(left, right, mono, stereo, n);
    return cudaGetLastError();
}

__global__ void k_pcm_pack(int16_t *pcm, const float *left, const float *right, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        pcm[2 * i] = (int16_t)pcm_s16(right[i]);
        pcm[2 * i + 1] = (int16_t)pcm_s16(left[i]);
    }
}

cudaError_t launch_pcm_pack(int16_t *pcm, const float *left, const float *right, size_t n,
                            cudaStream_t s)
{
    if (n == 0)
        return cudaSuccess;
    k_pcm_pack<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(pcm, left, right, n);
    return cudaGetLastError();
}

}  // namespace fmrx
