// EXPERIMENT, NOT BUILT (kept for the record; see DESIGN.md 4.1 "measured and rejected").
// Result on B200, 128 tracks x 5 min: 16.55 ms against 11.98 ms for stft_kernel.  The kernel was parity-green (all GPU tests),
// 85-93 % of the slots took the branch-free fused path, but with one CTA per SM there are only two warps per scheduler and
// the in-order issue of each warp exposes every dependency and shared-memory wait that sixteen warps hide in stft_kernel;
// instruction-level interleaving of two frames inside a thread does not make up for the lost thread-level parallelism.
// To build it again: #include it in tomatis_b200.cu after unit_epilogue and launch stft_dual_kernel<<<n_sms, 256, kDualSmem>>>.
//
// Dual-stream variant of stft_kernel (included inside the anonymous namespace of tomatis_b200.cu).
//
// One CTA per SM, 256 threads, every thread carries TWO frames of two different work units (streams X and Y) half a
// frame apart: while X is in its inner phase (stages B, C, C', B': warp-private E2 exchanges) Y is in its outer phase
// (stage A' + overlap-add tail of its previous frame, then gather + window + stage A of its next frame), and vice versa.
// The two phases of one slot are independent instruction streams of the same warp, so the shared-memory traffic of one
// stream is covered by the butterflies of the other *by construction* instead of by the luck of how two CTAs drift, and one
// CTA barrier per slot serves both streams (E1 of X and E1 of Y are published at the same point).
//
// Same building blocks as stft_kernel: fft4096.cuh stages, Park<0> (tensor-memory parking: each stream has its own 112
// columns per warp, 448 of the 512 columns per lane quarter), unit_epilogue (peaks + fused limiter), the global work queue.
#pragma once

struct DualCtx {                 // per stream, shared memory, written by thread 0 when a unit is claimed
    const float2* in_u;          // element 0 = sample 0 of the unit's first frame (thread offset not included)
    float2* out_u;
    const uint16_t* rows;        // track's row indices (index = frame)
    const TrackDev* trp;
    int b0, last;                // first output block, index of the unit's last frame (frames i = 0..last, f = b0-1+i)
    int n_frames, chunk;
    int in_lo, in_hi, out_lo, out_hi;
    int edge_lo, edge_hi;
    int valid;                   // 0: the queue is empty
    int pad_;
};

struct DualRegs {                // per stream, registers
    float2 v[16];
    float2 pf[8];
    float peak;
    int i;                       // frame in flight (index within the unit), -1 = none
    int have;                    // the frame in flight exists (0 <= f < n_frames)
    int do_pf;                   // pf holds half i+2, to be parked after stage B
    int unit;                    // a unit is active
};

constexpr int kDualSmem = 2 * (4096 + kExchFloat2) * (int)sizeof(float2) + 2 * (int)sizeof(DualCtx) + 128;

__device__ __forceinline__ void dual_load_half(const DualCtx& cx, int h, int t, float2 (&x)[8]) {
    const int p0 = h * kHop;
    const float2* src = cx.in_u + p0 + t;
    if (p0 >= cx.in_lo && p0 + kHop <= cx.in_hi) {
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = ld_stream(src + 256 * j);
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int p = p0 + 256 * j + t;
            x[j] = (p >= cx.in_lo && p < cx.in_hi) ? __ldg(src + 256 * j) : make_float2(0.f, 0.f);
        }
    }
}

// claim the next unit for a stream (all threads; contains CTA barriers)
__device__ __forceinline__ void dual_claim(const StftParams& prm, DualCtx* cx, DualRegs& r, const Park<0>& pk, int t) {
    __syncthreads();                                    // everybody is done with the previous contents of *cx
    if (t == 0) {
        const int u = atomicAdd(prm.unit_counter, 1);
        if (u < prm.n_units) {
            const UnitDev un = prm.units[u];
            const TrackDev* trp = prm.tracks + un.track;
            const long long upos = trp->first_start + (long long)(un.b0 - 1) * kHop;
            const long long span = (long long)(un.b1 - un.b0 + 2) * kHop + kNfft;
            cx->in_u = trp->in + (upos - trp->in_origin);
            cx->out_u = trp->out + (upos - trp->out_origin);
            cx->rows = prm.rows + trp->frame_base;
            cx->trp = trp;
            cx->b0 = un.b0;
            cx->last = un.b1 - un.b0;
            cx->n_frames = trp->n_frames;
            cx->chunk = un.chunk;
            cx->in_lo = (int)max(-span, min(span, trp->in_lo - upos));
            cx->in_hi = (int)max(-span, min(span, trp->in_hi - upos));
            cx->out_lo = (int)max(-span, min(span, trp->out_lo - upos));
            cx->out_hi = (int)max(-span, min(span, trp->out_hi - upos));
            cx->edge_lo = trp->edge_lo;
            cx->edge_hi = trp->edge_hi;
            cx->valid = 1;
        } else {
            cx->valid = 0;
        }
    }
    __syncthreads();
    r.unit = cx->valid;
    r.i = -1;
    r.have = 0;
    r.do_pf = 0;
    r.peak = 0.f;
    if (r.unit) {                                       // prologue: zero carry, halves 0 and 1 -> staging slots
        float2 x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = make_float2(0.f, 0.f);
        pk.store_carry(x, t);
        dual_load_half(*cx, 0, t, x);
        pk.stage_put(0, x);
        dual_load_half(*cx, 1, t, x);
        pk.stage_put(1, x);
        pk.sync_stores();
    }
}

// outer phase, second part: gather + window + stage A of the next frame of the unit -> E1 buffer
__device__ __forceinline__ void dual_begin(const DualCtx& cx, DualRegs& r, const Park<0>& pk, float2* bufP, int t) {
    r.i += 1;
    const int f = cx.b0 - 1 + r.i;
    r.have = (f >= 0) && (f < cx.n_frames);
    r.do_pf = r.i < cx.last;
    pk.sync_stores();
    if (r.have) {
        float fa[16], fb[16];
        pk.stage_get_windowed(r.i & 1, r.v, fa, fb);
        if (r.do_pf) dual_load_half(cx, r.i + 2, t, r.pf);
        dft16<false>(r.v);                                                        // A
        pk.twiddle_a_fwd(r.v, TwBase{}, fa, fb);
        st_e1a(r.v, t, bufP);
    } else if (r.do_pf) {
        dual_load_half(cx, r.i + 2, t, r.pf);
    }
}

// inner phase: stages B, C, gain, C', B' (E1 -> ... -> E1), warp-private E2 exchanges in between
__device__ __forceinline__ void dual_inner(const StftParams& prm, const DualCtx& cx, DualRegs& r, const Park<0>& pk,
                                           const TwBase wb, float2* bufP, float2* bufQ, int t) {
    if (r.i < 0) return;
    if (r.have) {
        const int row = cx.rows[cx.b0 - 1 + r.i];
        const float4* g4 = reinterpret_cast<const float4*>(prm.gperm + (size_t)row * kNfft + t * 16);
        const float4 g0 = __ldg(g4), g1 = __ldg(g4 + 1), g2 = __ldg(g4 + 2), g3 = __ldg(g4 + 3);
        ld_e1b(r.v, t, bufP);
        dft16<false>(r.v);                                                        // B
        tw_pow<false>(r.v, wb);
        st_e2b(r.v, t, bufQ);
        __syncwarp();
        if (r.do_pf) pk.stage_put(r.i & 1, r.pf);
        ld_e2c(r.v, t, bufQ);
        dft16<false>(r.v);                                                        // C
        r.v[0] = cscale(r.v[0], g0.x); r.v[1] = cscale(r.v[1], g0.y); r.v[2] = cscale(r.v[2], g0.z); r.v[3] = cscale(r.v[3], g0.w);
        r.v[4] = cscale(r.v[4], g1.x); r.v[5] = cscale(r.v[5], g1.y); r.v[6] = cscale(r.v[6], g1.z); r.v[7] = cscale(r.v[7], g1.w);
        r.v[8] = cscale(r.v[8], g2.x); r.v[9] = cscale(r.v[9], g2.y); r.v[10] = cscale(r.v[10], g2.z); r.v[11] = cscale(r.v[11], g2.w);
        r.v[12] = cscale(r.v[12], g3.x); r.v[13] = cscale(r.v[13], g3.y); r.v[14] = cscale(r.v[14], g3.z); r.v[15] = cscale(r.v[15], g3.w);
        dft16<true>(r.v);                                                         // C'
        st_e2c(r.v, t, bufQ);
        __syncwarp();
        ld_e2b(r.v, t, bufQ);
        tw_pow<true>(r.v, wb);
        dft16<true>(r.v);                                                         // B'
        st_e1b(r.v, t, bufP);
    } else if (r.do_pf) {
        pk.stage_put(r.i & 1, r.pf);
    }
}

// outer phase, first part: stage A' of the frame in flight, synthesis window, overlap-add, store; unit epilogue after the
// unit's last frame (all threads; the epilogue contains CTA barriers)
__device__ __forceinline__ void dual_finish(const StftParams& prm, const DualCtx& cx, DualRegs& r, const Park<0>& pk,
                                            float2* bufP, float* red, int t) {
    if (r.i < 0) return;
    float s[16];
    float2 c[8];
    if (r.have) {
        float ta[16], tb2[16], cr[16];
        pk.inv_fetch_issue(ta, tb2, s, cr, t);
        ld_e1a(r.v, t, bufP);
        pk.inv_fetch_apply(r.v, TwBase{}, ta, tb2, cr, c);
        dft16<true>(r.v);                                                         // A'
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) r.v[j] = make_float2(0.f, 0.f);
        pk.load_tail(s, c, t);
    }
    const int f = cx.b0 - 1 + r.i;
    const int rel = r.i * kHop;
    const bool edge_blk = (f == 0 && cx.edge_lo) || (f == cx.n_frames && cx.edge_hi);
    if (f >= cx.b0 && !edge_blk) {
        float2* dst = cx.out_u + rel + t;
        if ((rel >= cx.out_lo) && (rel + kHop <= cx.out_hi)) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float2 o = __ffma2_rn(r.v[j], make_float2(s[j], s[j]), c[j]);
                st_stream(dst + 256 * j, o);
                r.peak = fmaxf(r.peak, fmaxf(fabsf(o.x), fabsf(o.y)));
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float2 o = __ffma2_rn(r.v[j], make_float2(s[j], s[j]), c[j]);
                const int p = rel + 256 * j + t;
                if (p >= cx.out_lo && p < cx.out_hi) {
                    st_stream(dst + 256 * j, o);
                    r.peak = fmaxf(r.peak, fmaxf(fabsf(o.x), fabsf(o.y)));
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) c[j] = cscale(r.v[j + 8], s[j + 8]);
    pk.store_carry(c, t);
    if (r.i == cx.last) {
        unit_epilogue(prm, cx.chunk, cx.trp, r.peak, t, red);
        r.unit = 0;
        r.i = -1;
    }
}

__device__ __forceinline__ void dual_outer(const StftParams& prm, DualCtx* cx, DualRegs& r, const Park<0>& pk, float2* bufP,
                                           float* red, int t) {
    dual_finish(prm, *cx, r, pk, bufP, red, t);
    if (!r.unit && cx->valid) dual_claim(prm, cx, r, pk, t);      // cx->valid == 0: the queue ran dry earlier
    if (r.unit) dual_begin(*cx, r, pk, bufP, t);
}


// Steady-state slot, branch-free: stream I runs its inner phase while stream O finishes frame O.i and begins frame O.i+1.
// Both are straight-line code in ONE basic block, written stage by stage side by side, so the scheduler can cover the
// shared-memory and tensor-memory latencies of one stream with the butterflies of the other.
// Preconditions (checked by dual_fast_ok): I: frame in flight exists, next half to park; O: frame in flight exists, is neither
// the unit's warm-up nor its last frame, its block is a whole interior block of the output window, the next frame exists and
// its prefetch half lies inside the input window.
__device__ __forceinline__ bool dual_fast_ok(const DualCtx& ci, const DualRegs& I, const DualCtx& co, const DualRegs& O) {
    if (!(I.i >= 0 && I.have && I.do_pf)) return false;
    if (!(O.i >= 1 && O.have && O.i + 1 < co.last)) return false;
    const int f = co.b0 - 1 + O.i, rel = O.i * kHop;
    if ((f == 0 && co.edge_lo) || f + 1 >= co.n_frames) return false;
    if (!(rel >= co.out_lo && rel + kHop <= co.out_hi)) return false;
    const int p0 = (O.i + 3) * kHop;
    return p0 >= co.in_lo && p0 + kHop <= co.in_hi;
}

__device__ __forceinline__ void dual_slot_fast(const StftParams& prm, const DualCtx& ci, DualRegs& I, const Park<0>& pi,
                                               float2* bufPI, float2* bufQI, const DualCtx& co, DualRegs& O, const Park<0>& po,
                                               float2* bufPO, const TwBase wb, int t) {
    // ---- I: gain row + E1 read            | O: tail operands + E1 read
    const int row = ci.rows[ci.b0 - 1 + I.i];
    const float4* g4 = reinterpret_cast<const float4*>(prm.gperm + (size_t)row * kNfft + t * 16);
    const float4 g0 = __ldg(g4), g1 = __ldg(g4 + 1), g2 = __ldg(g4 + 2), g3 = __ldg(g4 + 3);
    float ta[16], tb2[16], cr[16], s[16];
    float2 c[8];
    po.inv_fetch_issue(ta, tb2, s, cr, t);
    ld_e1b(I.v, t, bufPI);
    ld_e1a(O.v, t, bufPO);
    // ---- I: stage B                        | O: stage A'
    dft16<false>(I.v);
    po.inv_fetch_apply(O.v, TwBase{}, ta, tb2, cr, c);
    tw_pow<false>(I.v, wb);
    dft16<true>(O.v);
    st_e2b(I.v, t, bufQI);
    __syncwarp();
    pi.stage_put(I.i & 1, I.pf);
    ld_e2c(I.v, t, bufQI);
    // ---- I: stage C, gain, C'              | O: synthesis window, overlap-add, store, carry
    {
        float2* dst = co.out_u + O.i * kHop + t;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float2 o = __ffma2_rn(O.v[j], make_float2(s[j], s[j]), c[j]);
            st_stream(dst + 256 * j, o);
            O.peak = fmaxf(O.peak, fmaxf(fabsf(o.x), fabsf(o.y)));
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) c[j] = cscale(O.v[j + 8], s[j + 8]);
        po.store_carry(c, t);
    }
    dft16<false>(I.v);
    I.v[0] = cscale(I.v[0], g0.x); I.v[1] = cscale(I.v[1], g0.y); I.v[2] = cscale(I.v[2], g0.z); I.v[3] = cscale(I.v[3], g0.w);
    I.v[4] = cscale(I.v[4], g1.x); I.v[5] = cscale(I.v[5], g1.y); I.v[6] = cscale(I.v[6], g1.z); I.v[7] = cscale(I.v[7], g1.w);
    I.v[8] = cscale(I.v[8], g2.x); I.v[9] = cscale(I.v[9], g2.y); I.v[10] = cscale(I.v[10], g2.z); I.v[11] = cscale(I.v[11], g2.w);
    I.v[12] = cscale(I.v[12], g3.x); I.v[13] = cscale(I.v[13], g3.y); I.v[14] = cscale(I.v[14], g3.z); I.v[15] = cscale(I.v[15], g3.w);
    dft16<true>(I.v);
    st_e2c(I.v, t, bufQI);
    __syncwarp();
    // ---- I: E2 read, stage B'              | O: next frame: gather from tensor memory, window, prefetch, stage A
    O.i += 1;
    {
        float fa[16], fb[16];
        po.sync_stores();
        po.stage_get_windowed(O.i & 1, O.v, fa, fb);
        const float2* src = co.in_u + (O.i + 2) * kHop + t;
#pragma unroll
        for (int j = 0; j < 8; ++j) O.pf[j] = ld_stream(src + 256 * j);
        ld_e2b(I.v, t, bufQI);
        dft16<false>(O.v);
        tw_pow<true>(I.v, wb);
        po.twiddle_a_fwd(O.v, TwBase{}, fa, fb);
        dft16<true>(I.v);
        st_e1a(O.v, t, bufPO);
        st_e1b(I.v, t, bufPI);
    }
}

__global__ void __launch_bounds__(kThreads, 1) stft_dual_kernel(const StftParams prm) {
    extern __shared__ __align__(16) unsigned char smraw[];
    float2* bufPX = reinterpret_cast<float2*>(smraw);
    float2* bufQX = bufPX + 4096;
    float2* bufPY = bufQX + kExchFloat2;
    float2* bufQY = bufPY + 4096;
    unsigned char* tail = reinterpret_cast<unsigned char*>(bufQY + kExchFloat2);
    float* red = reinterpret_cast<float*>(tail + 16);                  // 9 floats
    DualCtx* cxX = reinterpret_cast<DualCtx*>(tail + 64);
    DualCtx* cxY = cxX + 1;
    const int t = threadIdx.x, warp = t >> 5;

    // tensor memory: all 512 columns; per warp two stream regions of kTmemWarpCols
    uint32_t* slot = reinterpret_cast<uint32_t*>(tail);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (t == 0) { cxX->valid = 1; cxY->valid = 1; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    Park<0> pkX, pkY;
    pkX.base = *slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 2 * kTmemWarpCols);
    pkY.base = pkX.base + kTmemWarpCols;
    pkX.fill_tables(t, prm.win, prm.swin, prm.tw_a, prm.post_gain);
    pkY.fill_tables(t, prm.win, prm.swin, prm.tw_a, prm.post_gain);

    const float4* tb4 = reinterpret_cast<const float4*>(prm.tw_bases) + 2 * t;
    const float4 bb = __ldg(tb4 + 1);
    const TwBase wb = {make_float2(bb.x, bb.y), make_float2(bb.z, bb.w)};

    DualRegs X, Y;
    X.unit = Y.unit = 0;
    X.i = Y.i = -1;
    X.have = Y.have = X.do_pf = Y.do_pf = 0;
    X.peak = Y.peak = 0.f;
    // X starts a frame now, Y half a frame later
    dual_outer(prm, cxX, X, pkX, bufPX, red, t);
    __syncthreads();
    while (X.unit || Y.unit) {
        if (dual_fast_ok(*cxX, X, *cxY, Y)) {                            // slot alpha: X inner | Y outer
            dual_slot_fast(prm, *cxX, X, pkX, bufPX, bufQX, *cxY, Y, pkY, bufPY, wb, t);
        } else {
            dual_inner(prm, *cxX, X, pkX, wb, bufPX, bufQX, t);
            dual_outer(prm, cxY, Y, pkY, bufPY, red, t);
        }
        __syncthreads();
        if (dual_fast_ok(*cxY, Y, *cxX, X)) {                            // slot beta: Y inner | X outer
            dual_slot_fast(prm, *cxY, Y, pkY, bufPY, bufQY, *cxX, X, pkX, bufPX, wb, t);
        } else {
            dual_inner(prm, *cxY, Y, pkY, wb, bufPY, bufQY, t);
            dual_outer(prm, cxX, X, pkX, bufPX, red, t);
        }
        __syncthreads();
    }
    pkX.sync_stores();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(*slot) : "memory");
}
