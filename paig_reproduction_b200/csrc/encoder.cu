// ConvolutionalEncoder (blocks.py:52-103) and VelocityEncoder (blocks.py:8-49): forward and hand-written backward.
//
// The two UNets (ShallowUNet blocks.py:240-308 for H < 40, UNet blocks.py:106-237 otherwise) are described
// as data (layout.h: buffers + op list) and executed by one small interpreter, forward and in reverse.  The
// skip concats are channel slices of shared buffers, so "cat" costs nothing.  Every activation is kept in the
// workspace for the backward pass (440 KB per 32x32 frame; the step is compute-bound, not HBM-bound).
#include "common.cuh"
#include "internal.h"
#include "layout.h"

#include <cstdlib>

namespace paig {

// ---------------------------------------------------------------------------------------------------------
// UNet tables
// ---------------------------------------------------------------------------------------------------------
static void add_buf(UNetDesc& u, int C, int shift) { u.bufs[u.nbufs++] = BufDesc{C, shift}; }
static void add_op(UNetDesc& u, int kind, int layer, Ref in, Ref out, int relu) {
    u.ops[u.nops++] = Op{kind, layer, in, out, relu};
}

UNetDesc make_unet(int deep, int n) {
    UNetDesc u;
    const Ref X{-1, 0, 3}, LOG{-2, 0, n};
    if (!deep) {
        // blocks.py:278-308, hidden 8.  ReLU after every conv except c7 and c10; the 1x1 head keeps its ReLU (Q9).
        const int h = 8;
        enum { A1, CATB, P1, A3, CATA, P2, A5, A6, U1, A8, A9, U2, A11, A12 };
        add_buf(u, h, 0);       // A1   c1
        add_buf(u, 3 * h, 0);   // CATB [0:16) c10 | [16:24) x1 = c2
        add_buf(u, h, 1);       // P1
        add_buf(u, 2 * h, 1);   // A3   c3
        add_buf(u, 4 * h, 1);   // CATA [0:16) c7 | [16:32) x2 = c4
        add_buf(u, 2 * h, 2);   // P2
        add_buf(u, 4 * h, 2);   // A5   c5
        add_buf(u, 4 * h, 2);   // A6   c6
        add_buf(u, 4 * h, 1);   // U1
        add_buf(u, 2 * h, 1);   // A8   c8
        add_buf(u, 2 * h, 1);   // A9   c9
        add_buf(u, 2 * h, 0);   // U2
        add_buf(u, h, 0);       // A11  c11
        add_buf(u, h, 0);       // A12  c12
        add_op(u, OP_CONV, 0, X, Ref{A1, 0, h}, 1);
        add_op(u, OP_CONV, 1, Ref{A1, 0, h}, Ref{CATB, 2 * h, h}, 1);
        add_op(u, OP_POOL, -1, Ref{CATB, 2 * h, h}, Ref{P1, 0, h}, 0);
        add_op(u, OP_CONV, 2, Ref{P1, 0, h}, Ref{A3, 0, 2 * h}, 1);
        add_op(u, OP_CONV, 3, Ref{A3, 0, 2 * h}, Ref{CATA, 2 * h, 2 * h}, 1);
        add_op(u, OP_POOL, -1, Ref{CATA, 2 * h, 2 * h}, Ref{P2, 0, 2 * h}, 0);
        add_op(u, OP_CONV, 4, Ref{P2, 0, 2 * h}, Ref{A5, 0, 4 * h}, 1);
        add_op(u, OP_CONV, 5, Ref{A5, 0, 4 * h}, Ref{A6, 0, 4 * h}, 1);
        add_op(u, OP_UP, -1, Ref{A6, 0, 4 * h}, Ref{U1, 0, 4 * h}, 0);
        add_op(u, OP_CONV, 6, Ref{U1, 0, 4 * h}, Ref{CATA, 0, 2 * h}, 0);
        add_op(u, OP_CONV, 7, Ref{CATA, 0, 4 * h}, Ref{A8, 0, 2 * h}, 1);
        add_op(u, OP_CONV, 8, Ref{A8, 0, 2 * h}, Ref{A9, 0, 2 * h}, 1);
        add_op(u, OP_UP, -1, Ref{A9, 0, 2 * h}, Ref{U2, 0, 2 * h}, 0);
        add_op(u, OP_CONV, 9, Ref{U2, 0, 2 * h}, Ref{CATB, 0, 2 * h}, 0);
        add_op(u, OP_CONV, 10, Ref{CATB, 0, 3 * h}, Ref{A11, 0, h}, 1);
        add_op(u, OP_CONV, 11, Ref{A11, 0, h}, Ref{A12, 0, h}, 1);
        add_op(u, OP_HEAD, 12, Ref{A12, 0, h}, LOG, 1);
        u.logits_relu = 1;
    } else {
        // blocks.py:172-237, hidden 16.  No ReLU after c9, c12, c15 and the 1x1 head c18.
        const int h = 16;
        enum { A1, CATC, P1, A3, CATB, P2, A5, CATA, P3, A7, A8, U1, A10, A11, U2, A13, A14, U3, A16, A17 };
        add_buf(u, h, 0);        // A1   c1
        add_buf(u, 3 * h, 0);    // CATC [0:32) c15 | [32:48) x1 = c2
        add_buf(u, h, 1);        // P1
        add_buf(u, 2 * h, 1);    // A3   c3
        add_buf(u, 4 * h, 1);    // CATB [0:32) c12 | [32:64) x2 = c4
        add_buf(u, 2 * h, 2);    // P2
        add_buf(u, 4 * h, 2);    // A5   c5
        add_buf(u, 6 * h, 2);    // CATA [0:32) c9 | [32:96) x3 = c6
        add_buf(u, 4 * h, 3);    // P3
        add_buf(u, 8 * h, 3);    // A7   c7
        add_buf(u, 8 * h, 3);    // A8   c8
        add_buf(u, 8 * h, 2);    // U1
        add_buf(u, 4 * h, 2);    // A10  c10
        add_buf(u, 4 * h, 2);    // A11  c11
        add_buf(u, 4 * h, 1);    // U2
        add_buf(u, 2 * h, 1);    // A13  c13
        add_buf(u, 2 * h, 1);    // A14  c14
        add_buf(u, 2 * h, 0);    // U3
        add_buf(u, h, 0);        // A16  c16
        add_buf(u, h, 0);        // A17  c17
        add_op(u, OP_CONV, 0, X, Ref{A1, 0, h}, 1);
        add_op(u, OP_CONV, 1, Ref{A1, 0, h}, Ref{CATC, 2 * h, h}, 1);
        add_op(u, OP_POOL, -1, Ref{CATC, 2 * h, h}, Ref{P1, 0, h}, 0);
        add_op(u, OP_CONV, 2, Ref{P1, 0, h}, Ref{A3, 0, 2 * h}, 1);
        add_op(u, OP_CONV, 3, Ref{A3, 0, 2 * h}, Ref{CATB, 2 * h, 2 * h}, 1);
        add_op(u, OP_POOL, -1, Ref{CATB, 2 * h, 2 * h}, Ref{P2, 0, 2 * h}, 0);
        add_op(u, OP_CONV, 4, Ref{P2, 0, 2 * h}, Ref{A5, 0, 4 * h}, 1);
        add_op(u, OP_CONV, 5, Ref{A5, 0, 4 * h}, Ref{CATA, 2 * h, 4 * h}, 1);
        add_op(u, OP_POOL, -1, Ref{CATA, 2 * h, 4 * h}, Ref{P3, 0, 4 * h}, 0);
        add_op(u, OP_CONV, 6, Ref{P3, 0, 4 * h}, Ref{A7, 0, 8 * h}, 1);
        add_op(u, OP_CONV, 7, Ref{A7, 0, 8 * h}, Ref{A8, 0, 8 * h}, 1);
        add_op(u, OP_UP, -1, Ref{A8, 0, 8 * h}, Ref{U1, 0, 8 * h}, 0);
        add_op(u, OP_CONV, 8, Ref{U1, 0, 8 * h}, Ref{CATA, 0, 2 * h}, 0);
        add_op(u, OP_CONV, 9, Ref{CATA, 0, 6 * h}, Ref{A10, 0, 4 * h}, 1);
        add_op(u, OP_CONV, 10, Ref{A10, 0, 4 * h}, Ref{A11, 0, 4 * h}, 1);
        add_op(u, OP_UP, -1, Ref{A11, 0, 4 * h}, Ref{U2, 0, 4 * h}, 0);
        add_op(u, OP_CONV, 11, Ref{U2, 0, 4 * h}, Ref{CATB, 0, 2 * h}, 0);
        add_op(u, OP_CONV, 12, Ref{CATB, 0, 4 * h}, Ref{A13, 0, 2 * h}, 1);
        add_op(u, OP_CONV, 13, Ref{A13, 0, 2 * h}, Ref{A14, 0, 2 * h}, 1);
        add_op(u, OP_UP, -1, Ref{A14, 0, 2 * h}, Ref{U3, 0, 2 * h}, 0);
        add_op(u, OP_CONV, 14, Ref{U3, 0, 2 * h}, Ref{CATC, 0, 2 * h}, 0);
        add_op(u, OP_CONV, 15, Ref{CATC, 0, 3 * h}, Ref{A16, 0, h}, 1);
        add_op(u, OP_CONV, 16, Ref{A16, 0, h}, Ref{A17, 0, h}, 1);
        add_op(u, OP_HEAD, 17, Ref{A17, 0, h}, LOG, 0);
        u.logits_relu = 0;
    }
    return u;
}

Layout make_layout(const paig_task* t, int B) {
    Layout L;
    L.d = dims_of(t);
    const Dims& d = L.d;
    L.B = B;
    L.N = B * d.e;
    L.unet = make_unet(t->deep_unet, d.n);
    L.K = t->deep_unet ? 3 * (d.H / 2) * (d.H / 2) : 3 * d.HW;
    size_t off = 0;
    auto take = [&](size_t floats) {
        size_t o = off;
        off += align64(floats);
        return o;
    };
    const size_t N = (size_t)L.N, nN = (size_t)d.n * L.N, nB = (size_t)d.n * B;
    // PAIG_FLAG_INFERENCE (forward only: eval_performance, base.py:174-218): nothing is kept for a backward pass.  With
    // the ShallowUNet the fused forward kernel then stores no activation at all, and every gradient / staging region
    // has size zero: ~0.65 MB per sequence instead of ~14.5 MB (spring_color_half test mode), so the 8192-sequence
    // evaluation sweep needs ~5 GB, not 119 GB.
    const bool inf = (t->flags & PAIG_FLAG_INFERENCE) != 0;
    static const bool layerwise = getenv("PAIG_UNET_LAYERWISE") != nullptr;
    const bool acts_on_chip = inf && !t->deep_unet && !layerwise && d.H <= 36;
    auto bwd = [&](size_t floats) { return inf ? (size_t)0 : floats; };     // regions only a backward pass touches
    size_t max_w = 0, sum_w = 0;
    for (int i = 0; i < L.unet.nbufs; ++i) {
        const int S = d.H >> L.unet.bufs[i].shift;
        const size_t fl = N * L.unet.bufs[i].C * S * S;
        L.act[i] = take(acts_on_chip ? 0 : fl);
        L.grad[i] = take(bwd(fl));
    }
    for (int i = 0; i < L.unet.nops; ++i) {
        const Op& op = L.unet.ops[i];
        if (op.kind == OP_CONV) {
            const size_t w = wgrad_partials_floats(op.in.C, op.out.C);
            max_w = w > max_w ? w : max_w;
            sum_w += align64(w);
        }
    }
    const size_t head = (size_t)592 * (kMaxObjs * 17);
    max_w = head > max_w ? head : max_w;
    sum_w += align64(head);                 // every layer keeps its own partials when the folds are batched
    max_w = sum_w > max_w ? sum_w : max_w;
    if (inf) max_w = 0;
    L.logits = take(N * d.n * d.HW);
    L.d_logits = take(bwd(N * d.n * d.HW));
    L.masks = take(N * (d.n + 1) * d.HW);
    L.A = take(nN * L.K);
    L.dA = take(bwd(nN * L.K));
    L.H1 = take(nN * kHidden); L.H2 = take(nN * kHidden); L.O3 = take(nN * 2);
    L.dH1 = take(bwd(nN * kHidden)); L.dH2 = take(bwd(nN * kHidden)); L.dO3 = take(bwd(nN * 2));
    L.enc_pos = take(N * 2 * d.n);
    L.d_enc_pos = take(bwd(N * 2 * d.n));
    const int vin = 2 * d.in;
    L.vin = take(nB * vin); L.v1 = take(nB * kVelHidden); L.v2 = take(nB * kVelHidden); L.vout = take(nB * 2);
    L.dvin = take(bwd(nB * vin)); L.dv1 = take(bwd(nB * kVelHidden)); L.dv2 = take(bwd(nB * kVelHidden)); L.dvout = take(bwd(nB * 2));
    const size_t seq = (size_t)B * (d.steps + 1) * 4 * d.n;
    L.seq = take(seq);
    L.d_seq = take(bwd(seq));
    L.d_state0 = take(bwd((size_t)B * 4 * d.n));
    const size_t CN = (size_t)d.n * d.t * d.t * 4 + 3 * d.HW;
    L.raw = take(CN); L.consts = take(CN); L.hidden = take(3 * kHidden); L.d_consts = take(bwd(CN));
    L.dec_partials = take(bwd(decode_partials_floats(t)));
    L.tmpl_scratch = take(bwd(templates_scratch_floats(t)));
    L.sse = take((size_t)B * (d.e + d.steps));
    L.scales = take(d.e + d.steps);
    L.losses = take(8);
    L.dphys = take(bwd(2 * rollout_scratch_doubles(B)));   // arrival counter + per-block fp64 partials of the rollout backward
    {   // room for the split-K partials of encoder.l1 forward ([splits][nN][200]) and weight gradient ([splits][200][K])
        const size_t fwd = (size_t)cdiv(L.K, 256) * nN * kHidden, wg = bwd(4 * (size_t)kHidden * L.K);
        const size_t sk = fwd > wg ? fwd : wg;
        max_w = sk > max_w ? sk : max_w;
    }
    L.partials = take(max_w);
    L.partials_floats = max_w;
    L.partials2_floats = bwd(4 * (size_t)kHidden * L.K);      // encoder.l1 weight gradient's split-K partials: it runs on
    L.partials2 = take(L.partials2_floats);                   // a side stream beside the UNet backward (own scratch)
    L.tc_scratch = take(bwd((size_t)L.K * kHidden + (size_t)kHidden * nN + (size_t)L.K * nN));
    {
        size_t mx = 64;
        if (t->deep_unet)
            for (int i = 0; i < L.unet.nops; ++i)
                if (L.unet.ops[i].kind == OP_CONV) {
                    const size_t w = conv_tc_scratch_floats(L.unet.ops[i].in.C, L.unet.ops[i].out.C);
                    mx = w > mx ? w : mx;
                }
        L.convtc = take(mx);
    }
    L.wpack = take(unet_wpack_floats(L.unet, t));
    L.wpack_tc = take(t->deep_unet ? 0 : unet_tc_wpack_floats(L.unet));
    L.frames = take(acts_on_chip ? 0 : N * d.CHW);
    L.x_stage = take(bwd((size_t)B * d.T * d.CHW));
    L.x_stage2 = take(bwd((size_t)B * d.T * d.CHW));
    L.total = off;
    return L;
}

// ---------------------------------------------------------------------------------------------------------
// UNet interpreter
// ---------------------------------------------------------------------------------------------------------
struct View {
    float* p;
    long bs;
    int S;
};

static View view_of(const Layout& L, float* ws, const Ref& r, bool grad) {
    View v;
    const BufDesc& b = L.unet.bufs[r.buf];
    v.S = L.d.H >> b.shift;
    v.bs = (long)b.C * v.S * v.S;
    v.p = ws + (grad ? L.grad[r.buf] : L.act[r.buf]) + (long)r.c0 * v.S * v.S;
    return v;
}

static int unet_forward(const paig_task* t, const paig_params* p, const Layout& L, float* ws, const float* frames,
                        cudaStream_t st) {
    const Dims& d = L.d;
    for (int i = 0; i < L.unet.nops; ++i) {
        const Op& op = L.unet.ops[i];
        int rc = 0;
        if (op.kind == OP_CONV) {
            ConvArgs a;
            if (op.in.buf == -1) {
                a.in = frames; a.in_bs = d.CHW; a.S = d.H;
            } else {
                View v = view_of(L, ws, op.in, false);
                a.in = v.p; a.in_bs = v.bs; a.S = v.S;
            }
            View o = view_of(L, ws, op.out, false);
            a.Cin = op.in.C;
            a.w = p->conv[op.layer].w; a.b = p->conv[op.layer].b;
            a.out = o.p; a.out_bs = o.bs; a.Cout = op.out.C;
            a.N = L.N; a.relu = op.relu;
            a.tc_scratch = t->deep_unet ? ws + L.convtc : nullptr;
            rc = conv3x3(a, st);
        } else if (op.kind == OP_POOL) {
            View v = view_of(L, ws, op.in, false), o = view_of(L, ws, op.out, false);
            rc = maxpool2(v.p, v.bs, o.p, o.bs, op.in.C, o.S, L.N, st);
        } else if (op.kind == OP_UP) {
            View v = view_of(L, ws, op.in, false), o = view_of(L, ws, op.out, false);
            rc = upsample2(v.p, v.bs, o.p, o.bs, op.in.C, v.S, L.N, st);
        } else {
            View v = view_of(L, ws, op.in, false);
            rc = conv1x1_forward(v.p, v.bs, op.in.C, p->conv[op.layer].w, p->conv[op.layer].b, ws + L.logits,
                                 (long)d.n * d.HW, d.n, d.H, L.N, op.relu, st);
        }
        if (rc) return rc;
    }
    (void)t;
    return 0;
}

static int unet_backward(const paig_task* t, const paig_params* p, const paig_params* g, const Layout& L, float* ws,
                         const float* frames, cudaStream_t st) {
    const Dims& d = L.d;
    float* partials = ws + L.partials;
    // Data gradients of the whole UNet in one persistent kernel (unet_fused.cu) when it fits on chip: every conv
    // output's ReLU-gated gradient lands in the workspace; what remains per layer is its weight gradient.
    static const bool layerwise = getenv("PAIG_UNET_LAYERWISE") != nullptr;
    int frc = layerwise ? -1 : unet_tc_backward(t, p, L, ws, st);            // tcgen05 (32-px frames)
    if (frc < 0 && !layerwise) frc = unet_fused_backward(t, p, L, ws, st);
    if (frc > 0) return frc;
    if (frc == 0) {
        ReduceBatch folds;                               // one launch folds every layer's per-CTA partials at the end
        size_t poff = 0;
        // the layers' weight gradients are independent of one another: dealt over three streams, the tail wave of one
        // launch overlaps the first wave of the next (each layer keeps its own partials region)
        Side* sd = side_cur();
        cudaStream_t lanes[3] = {st, sd ? sd->s1 : st, sd ? sd->s2 : st};
        if (sd) { sd->fork1(); sd->fork2(); }
        int lane = 0;
        for (int i = L.unet.nops - 1; i >= 0; --i) {
            const Op& op = L.unet.ops[i];
            int rc = 0;
            cudaStream_t ls = lanes[lane % 3];
            if (op.kind == OP_HEAD) {
                View v = view_of(L, ws, op.in, false);
                rc = conv1x1_backward(v.p, v.bs, op.in.C, p->conv[op.layer].w, ws + L.d_logits, (long)d.n * d.HW,
                                      op.relu ? ws + L.logits : nullptr, (long)d.n * d.HW, d.n, d.H, L.N, nullptr, 0,
                                      g->conv[op.layer].w, g->conv[op.layer].b, partials + poff, ls, &folds);
                poff += align64((size_t)592 * (kMaxObjs * 17));
                ++lane;
            } else if (op.kind == OP_CONV) {
                View o = view_of(L, ws, op.out, false), go = view_of(L, ws, op.out, true);
                WgradArgs w;
                if (op.in.buf == -1) {
                    w.in = frames; w.in_bs = d.CHW;
                } else {
                    View v = view_of(L, ws, op.in, false);
                    w.in = v.p; w.in_bs = v.bs;
                }
                w.Cin = op.in.C;
                w.g = go.p; w.g_bs = go.bs; w.Cout = op.out.C;       // already gated by the layer's ReLU
                w.act = nullptr;
                w.S = o.S; w.N = L.N; w.partials = partials + poff;
                w.defer = &folds;
                poff += align64(wgrad_partials_floats(op.in.C, op.out.C));
                rc = conv3x3_wgrad(w, g->conv[op.layer].w, g->conv[op.layer].b, ls);
                ++lane;
            }
            if (rc) return rc;
        }
        if (sd) { sd->join1(); sd->join2(); }
        return reduce_partials_batch(folds, st);
    }
    for (int i = L.unet.nops - 1; i >= 0; --i) {
        const Op& op = L.unet.ops[i];
        int rc = 0;
        if (op.kind == OP_HEAD) {
            View v = view_of(L, ws, op.in, false), dv = view_of(L, ws, op.in, true);
            rc = conv1x1_backward(v.p, v.bs, op.in.C, p->conv[op.layer].w, ws + L.d_logits, (long)d.n * d.HW,
                                  op.relu ? ws + L.logits : nullptr, (long)d.n * d.HW, d.n, d.H, L.N, dv.p, dv.bs,
                                  g->conv[op.layer].w, g->conv[op.layer].b, partials, st);
        } else if (op.kind == OP_CONV) {
            View o = view_of(L, ws, op.out, false), go = view_of(L, ws, op.out, true);
            WgradArgs w;
            if (op.in.buf == -1) {
                w.in = frames; w.in_bs = d.CHW;
            } else {
                View v = view_of(L, ws, op.in, false);
                w.in = v.p; w.in_bs = v.bs;
            }
            w.Cin = op.in.C;
            w.g = go.p; w.g_bs = go.bs; w.Cout = op.out.C;
            // gate the gradient once in place: the TMA weight-gradient kernel needs it pre-gated, and the data-gradient
            // kernel reads it without a mask (which also lets it take the cp.async double-buffered variant)
            if (op.relu && (rc = relu_gate(go.p, go.bs, o.p, o.bs, op.out.C, o.S, L.N, st))) return rc;
            w.act = nullptr;
            w.S = o.S; w.N = L.N; w.partials = partials;
            rc = conv3x3_wgrad(w, g->conv[op.layer].w, g->conv[op.layer].b, st);
            if (!rc && op.in.buf >= 0) {        // no gradient w.r.t. the input frames
                View gi = view_of(L, ws, op.in, true);
                ConvArgs a;
                a.in = go.p; a.in_bs = go.bs; a.Cin = op.out.C;
                a.mask = nullptr; a.mask_bs = 0;
                a.w = p->conv[op.layer].w; a.transposed = 1;
                a.out = gi.p; a.out_bs = gi.bs; a.Cout = op.in.C;
                a.S = o.S; a.N = L.N;
                a.tc_scratch = t->deep_unet ? ws + L.convtc : nullptr;
                rc = conv3x3(a, st);
            }
        } else if (op.kind == OP_POOL) {
            View v = view_of(L, ws, op.in, false), gi = view_of(L, ws, op.in, true);
            View go = view_of(L, ws, op.out, true);
            rc = maxpool2_backward(v.p, v.bs, go.p, go.bs, gi.p, gi.bs, op.in.C, go.S, L.N, st);
        } else {
            View gi = view_of(L, ws, op.in, true), go = view_of(L, ws, op.out, true);
            rc = upsample2_backward(go.p, go.bs, gi.p, gi.bs, op.in.C, gi.S, L.N, st);
        }
        if (rc) return rc;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// mask softmax + masked objects (blocks.py:84-96) and the position head (blocks.py:101-102)
// ---------------------------------------------------------------------------------------------------------
// One thread per pixel (POOL: per 2x2 block).  A[o*N + f][c*KP + pp] = m_o * x_c  (POOL: 2x2 mean of it).
template <int NOBJ, bool POOL>
__global__ void __launch_bounds__(256) enc_masks_kernel(const float* __restrict__ logits, const float* __restrict__ x,
                                                        long x_seq_stride, int fps, int H, int N,
                                                        float* __restrict__ masks, float* __restrict__ A,
                                                        float* __restrict__ masked_objs) {
    const int HW = H * H;
    const int W2 = POOL ? H / 2 : H, KP = POOL ? HW / 4 : HW;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)N * KP) return;
    const int f = (int)(idx / KP), pp = (int)(idx % KP);
    const float* xf = x + (long)(f / fps) * x_seq_stride + (long)(f % fps) * 3 * HW;
    float accA[NOBJ][3];
#pragma unroll
    for (int o = 0; o < NOBJ; ++o)
#pragma unroll
        for (int c = 0; c < 3; ++c) accA[o][c] = 0.f;
    const int np = POOL ? 4 : 1;
    for (int s = 0; s < np; ++s) {
        const int p = POOL ? ((pp / W2) * 2 + (s >> 1)) * H + (pp % W2) * 2 + (s & 1) : pp;
        float z[NOBJ], m = 1.f;
#pragma unroll
        for (int o = 0; o < NOBJ; ++o) {
            z[o] = logits[((long)f * NOBJ + o) * HW + p];
            m = fmaxf(m, z[o]);
        }
        float den = 0.f, e[NOBJ + 1];
#pragma unroll
        for (int o = 0; o < NOBJ; ++o) {
            e[o] = expf(z[o] - m);
            den += e[o];
        }
        e[NOBJ] = expf(1.f - m);           // background logit: the constant 1 (blocks.py:86)
        den += e[NOBJ];
        const float inv = 1.f / den;
#pragma unroll
        for (int o = 0; o <= NOBJ; ++o) masks[((long)f * (NOBJ + 1) + o) * HW + p] = e[o] * inv;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float xv = xf[c * HW + p];
#pragma unroll
            for (int o = 0; o < NOBJ; ++o) {
                const float v = e[o] * inv * xv;
                accA[o][c] += v;
                if (masked_objs) masked_objs[(((long)o * N + f) * 3 + c) * HW + p] = v;
            }
        }
    }
#pragma unroll
    for (int o = 0; o < NOBJ; ++o)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            A[((long)o * N + f) * (3 * KP) + c * KP + pp] = POOL ? 0.25f * accA[o][c] : accA[o][c];
}

// d logits from dA:  dm_o = sum_c dA * x_c (x .25 under the 2x2 mean);  dz_k = m_k (dm_k - sum_j m_j dm_j)
template <int NOBJ, bool POOL>
__global__ void __launch_bounds__(256) enc_masks_bwd_kernel(const float* __restrict__ masks, const float* __restrict__ x,
                                                            long x_seq_stride, int fps, int H, int N,
                                                            const float* __restrict__ dA, float* __restrict__ d_logits) {
    const int HW = H * H;
    const int W2 = POOL ? H / 2 : H, KP = POOL ? HW / 4 : HW;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)N * HW) return;
    const int f = (int)(idx / HW), p = (int)(idx % HW);
    const int pp = POOL ? ((p / H) / 2) * W2 + (p % H) / 2 : p;
    const float* xf = x + (long)(f / fps) * x_seq_stride + (long)(f % fps) * 3 * HW;
    float dm[NOBJ], mk[NOBJ], dot = 0.f;
#pragma unroll
    for (int o = 0; o < NOBJ; ++o) {
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) s += dA[((long)o * N + f) * (3 * KP) + c * KP + pp] * xf[c * HW + p];
        dm[o] = POOL ? 0.25f * s : s;
        mk[o] = masks[((long)f * (NOBJ + 1) + o) * HW + p];
        dot += mk[o] * dm[o];
    }
#pragma unroll
    for (int o = 0; o < NOBJ; ++o) d_logits[((long)f * NOBJ + o) * HW + p] = mk[o] * (dm[o] - dot);
}

// enc_pos[f, 2o+c] = tanh(O3[o*N+f, c]) * H/2 + H/2      (blocks.py:101-102)
__global__ void __launch_bounds__(256) pos_head_kernel(const float* __restrict__ O3, int N, int n, float half,
                                                       float* __restrict__ enc_pos, float* __restrict__ enc_pos2) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * 2 * n) return;
    const int f = idx / (2 * n), k = idx % (2 * n), o = k >> 1, c = k & 1;
    const float v = tanhf(O3[((long)o * N + f) * 2 + c]) * half + half;
    enc_pos[idx] = v;
    if (enc_pos2) enc_pos2[idx] = v;
}
__global__ void __launch_bounds__(256) pos_head_bwd_kernel(const float* __restrict__ O3, int N, int n, float half,
                                                           const float* __restrict__ d_enc_pos,
                                                           float* __restrict__ dO3) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * 2 * n) return;
    const int f = idx / (2 * n), k = idx % (2 * n), o = k >> 1, c = k & 1;
    const long j = ((long)o * N + f) * 2 + c;
    const float th = tanhf(O3[j]);
    dO3[j] = d_enc_pos[idx] * half * (1.f - th * th);
}

template <int NOBJ>
static int masks_fwd(const Layout& L, bool pool, float* ws, const float* x, long ss, int fps, float* masks,
                     float* masked, cudaStream_t st) {
    const Dims& d = L.d;
    const long total = (long)L.N * (pool ? d.HW / 4 : d.HW);
    if (pool)
        launch(enc_masks_kernel<NOBJ, true>, dim3(cdiv(total, 256)), dim3(256), 0, st, (const float*)(ws + L.logits), x,
               ss, fps, d.H, L.N, masks, ws + L.A, masked);
    else
        launch(enc_masks_kernel<NOBJ, false>, dim3(cdiv(total, 256)), dim3(256), 0, st, (const float*)(ws + L.logits), x,
               ss, fps, d.H, L.N, masks, ws + L.A, masked);
    return check_launch("enc_masks");
}
template <int NOBJ>
static int masks_bwd(const Layout& L, bool pool, float* ws, const float* x, long ss, int fps, cudaStream_t st) {
    const Dims& d = L.d;
    const long total = (long)L.N * d.HW;
    if (pool)
        launch(enc_masks_bwd_kernel<NOBJ, true>, dim3(cdiv(total, 256)), dim3(256), 0, st, (const float*)(ws + L.masks),
               x, ss, fps, d.H, L.N, (const float*)(ws + L.dA), ws + L.d_logits);
    else
        launch(enc_masks_bwd_kernel<NOBJ, false>, dim3(cdiv(total, 256)), dim3(256), 0, st,
               (const float*)(ws + L.masks), x, ss, fps, d.H, L.N, (const float*)(ws + L.dA), ws + L.d_logits);
    return check_launch("enc_masks_bwd");
}

// frames of the encoder batch made contiguous [N,3,H,H]
__global__ void __launch_bounds__(256) gather_frames_kernel(const float4* __restrict__ x, long seq_stride4, int fps,
                                                            int chw4, long total4, float4* __restrict__ out) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total4) return;
    const long f = idx / chw4, r = idx % chw4;
    out[idx] = x[(f / fps) * seq_stride4 + (f % fps) * chw4 + r];
}

// ---- fused tail of the localisation MLP: relu(l2) -> l3 -> tanh * H/2 + H/2  (blocks.py:99-102) ----------------------
// After the big l1 product the rest of the MLP is small (M x 200 x 200 and M x 200 x 2) and was three launches forward
// and nine backward.  Forward: 16 rows per CTA, W2 transposed into shared memory (lane j = unit j).  Backward: 16 rows per
// CTA write their share of dW2 / dW3 / db2 / db3 to a partial block (folded in fixed order by reduce_partials_batch)
// and dH1 (gated by l1's ReLU) for the l1 backward GEMMs.
constexpr int kTailRows = 16;
// W2 lives in shared memory as [unit j][k] with rows padded to 204 floats: a row starts on a 16-byte boundary and the
// 16-byte loads of 8 consecutive units (row pitch 816 B = 12 banks mod 32) touch 8 distinct bank quads.  Both kernels keep
// the summation order of the scalar loops they replaced (k, r, j ascending), so results did not change bit for bit; what
// changed is the shared-memory traffic: 16-byte operand loads and register tiles instead of one 4-byte load per FMA
// (forward 40 -> ? us, backward 72 -> ? us at M = 2000; both were bound by the LSU pipe and by 4-byte global loads of W2).
constexpr int kTailPitch = 204;

// Loads are issued in batches of 8 (W2) / 4 (activation rows) per thread before the first store: one load per trip left the
// ~700-cycle L2 latency exposed 40 times over (W2 alone: 14 us of the 37 us forward kernel).
__device__ __forceinline__ void tail_load_w2(float* sW2, const float* __restrict__ W2, int tid) {
    constexpr int HID = kHidden, NV = HID * HID / 4;
    if ((reinterpret_cast<uintptr_t>(W2) & 15) == 0) {
        const float4* src = reinterpret_cast<const float4*>(W2);
        for (int b = 0; b < NV; b += 8 * 256) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = b + u * 256 + tid;
                if (i < NV) v[u] = src[i];
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = b + u * 256 + tid;
                if (i < NV) {
                    const int j = i / (HID / 4), c = i - j * (HID / 4);
                    *reinterpret_cast<float4*>(sW2 + j * kTailPitch + 4 * c) = v[u];
                }
            }
        }
    } else {
        for (int i = tid; i < HID * HID; i += 256) sW2[(i / HID) * kTailPitch + i % HID] = W2[i];
    }
}
// kTailRows rows of a [M][200] activation (rows past M read as zero) into a dense [16][200] shared array
__device__ __forceinline__ void tail_load_rows(float* dst, const float* __restrict__ src, int r0, int M, int tid) {
    constexpr int HID = kHidden, R = kTailRows, NV = R * HID / 4;
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = u * 256 + tid, r = i / (HID / 4);
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < NV && r0 + r < M) v[u] = reinterpret_cast<const float4*>(src + (long)r0 * HID)[i];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = u * 256 + tid;
            if (i < NV) reinterpret_cast<float4*>(dst)[i] = v[u];
        }
    } else {
        for (int i = tid; i < R * HID; i += 256) dst[i] = r0 + i / HID < M ? src[(long)(r0 + i / HID) * HID + i % HID] : 0.f;
    }
}

__global__ void __launch_bounds__(256) enc_tail_fwd_kernel(const float* __restrict__ H1, int M, int N, int n, float half,
                                                           const float* __restrict__ W2, const float* __restrict__ b2,
                                                           const float* __restrict__ W3, const float* __restrict__ b3,
                                                           float* __restrict__ H2, float* __restrict__ O3,
                                                           float* __restrict__ enc_pos, float* __restrict__ enc_pos2) {
    constexpr int HID = kHidden, R = kTailRows;
    PAIG_DYN_SMEM(float, smem);
    float* sW2 = smem;                                    // [j][k], rows of kTailPitch
    float (*sH1)[HID] = reinterpret_cast<float (*)[HID]>(sW2 + HID * kTailPitch);
    float (*sH2)[HID] = reinterpret_cast<float (*)[HID]>(&sH1[R][0]);
    const int tid = threadIdx.x, r0 = blockIdx.x * R;
    tail_load_w2(sW2, W2, tid);
    tail_load_rows(&sH1[0][0], H1, r0, M, tid);
    __syncthreads();
    if (tid < HID) {
        float acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = 0.f;
        const float* wrow = sW2 + tid * kTailPitch;
        for (int k = 0; k < HID; k += 4) {
            const float4 w = *reinterpret_cast<const float4*>(wrow + k);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float4 h = *reinterpret_cast<const float4*>(&sH1[r][k]);      // warp-wide broadcast
                acc[r] += h.x * w.x;
                acc[r] += h.y * w.y;
                acc[r] += h.z * w.z;
                acc[r] += h.w * w.w;
            }
        }
        const float bb = b2[tid];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float h = fmaxf(acc[r] + bb, 0.f);
            sH2[r][tid] = h;
            if (r0 + r < M) H2[(long)(r0 + r) * HID + tid] = h;
        }
    }
    __syncthreads();
    if (tid < R * 2) {
        const int r = tid >> 1, c = tid & 1, row = r0 + r;
        if (row < M) {
            float s = 0.f;
            for (int k = 0; k < HID; ++k) s += sH2[r][k] * W3[c * HID + k];
            s += b3[c];
            O3[(long)row * 2 + c] = s;
            const int o = row / N, f = row % N;                       // rows are object-major (blocks.py:88-93)
            const float v = tanhf(s) * half + half;
            enc_pos[(long)f * 2 * n + 2 * o + c] = v;
            if (enc_pos2) enc_pos2[(long)f * 2 * n + 2 * o + c] = v;
        }
    }
}

__global__ void __launch_bounds__(256) enc_tail_bwd_kernel(const float* __restrict__ d_enc_pos, int M, int N, int n, float half,
                                                           const float* __restrict__ W2, const float* __restrict__ W3,
                                                           const float* __restrict__ H1, const float* __restrict__ H2,
                                                           const float* __restrict__ O3, float* __restrict__ dH1,
                                                           float* __restrict__ partials, int stride) {
    constexpr int HID = kHidden, R = kTailRows;
    PAIG_DYN_SMEM(float, smem);
    float* sW2 = smem;                                    // [j][k], rows of kTailPitch
    float (*sH1)[HID] = reinterpret_cast<float (*)[HID]>(sW2 + HID * kTailPitch);
    float (*sH2)[HID] = reinterpret_cast<float (*)[HID]>(&sH1[R][0]);
    float (*sDz2)[HID] = reinterpret_cast<float (*)[HID]>(&sH2[R][0]);
    float (*sDo3)[2] = reinterpret_cast<float (*)[2]>(&sDz2[R][0]);
    const int tid = threadIdx.x, r0 = blockIdx.x * R;
    tail_load_w2(sW2, W2, tid);
    tail_load_rows(&sH1[0][0], H1, r0, M, tid);
    tail_load_rows(&sH2[0][0], H2, r0, M, tid);
    if (tid < R * 2) {
        const int r = tid >> 1, c = tid & 1, row = r0 + r;
        float v = 0.f;
        if (row < M) {
            const int o = row / N, f = row % N;
            const float th = tanhf(O3[(long)row * 2 + c]);
            v = d_enc_pos[(long)f * 2 * n + 2 * o + c] * half * (1.f - th * th);
        }
        sDo3[r][c] = v;
    }
    __syncthreads();
    float* out = partials + (size_t)blockIdx.x * stride;          // [dW2 | dW3 | db2 | db3]
    float* oW3 = out + HID * HID, *ob2 = oW3 + 2 * HID, *ob3 = ob2 + HID;
    // dZ2 = (W3^T dO3) gated by l2's ReLU
    for (int i = tid; i < R * HID; i += 256) {
        const int r = i / HID, k = i % HID;
        const float d = W3[k] * sDo3[r][0] + W3[HID + k] * sDo3[r][1];
        sDz2[r][k] = sH2[r][k] > 0.f ? d : 0.f;
    }
    for (int i = tid; i < 2 * HID; i += 256) {
        const int c = i / HID, k = i % HID;
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) s += sDo3[r][c] * sH2[r][k];
        oW3[i] = s;
    }
    if (tid < 2) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) s += sDo3[r][tid];
        ob3[tid] = s;
    }
    __syncthreads();
    // this chunk's share of dW2[j][k] = sum_r dZ2[r][j] H1[r][k]: thread = (20 units j) x (4 columns k), two passes
    if (tid < 250) {
        const int k4 = tid % 50, jg = tid / 50;
        for (int pass = 0; pass < 2; ++pass) {
            const int j0 = jg * 40 + pass * 20;
            float4 acc[20];
#pragma unroll
            for (int jj = 0; jj < 20; ++jj) acc[jj] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
            for (int r = 0; r < R; ++r) {
                const float4 h = *reinterpret_cast<const float4*>(&sH1[r][4 * k4]);
#pragma unroll
                for (int jj = 0; jj < 20; ++jj) {
                    const float dz = sDz2[r][j0 + jj];
                    acc[jj].x += dz * h.x;
                    acc[jj].y += dz * h.y;
                    acc[jj].z += dz * h.z;
                    acc[jj].w += dz * h.w;
                }
            }
#pragma unroll
            for (int jj = 0; jj < 20; ++jj) {
                float* o = out + (j0 + jj) * HID + 4 * k4;
                o[0] = acc[jj].x; o[1] = acc[jj].y; o[2] = acc[jj].z; o[3] = acc[jj].w;
            }
        }
    }
    if (tid < HID) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) s += sDz2[r][tid];
        ob2[tid] = s;
        // dH1 = (dZ2 W2) gated by l1's ReLU: thread = column k, all 16 rows in registers
        float acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = 0.f;
        for (int j = 0; j < HID; j += 4) {
            const float w0 = sW2[(j + 0) * kTailPitch + tid], w1 = sW2[(j + 1) * kTailPitch + tid];
            const float w2 = sW2[(j + 2) * kTailPitch + tid], w3 = sW2[(j + 3) * kTailPitch + tid];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float4 dz = *reinterpret_cast<const float4*>(&sDz2[r][j]);    // warp-wide broadcast
                acc[r] += w0 * dz.x;
                acc[r] += w1 * dz.y;
                acc[r] += w2 * dz.z;
                acc[r] += w3 * dz.w;
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (r0 + r < M) dH1[(long)(r0 + r) * HID + tid] = sH1[r][tid] > 0.f ? acc[r] : 0.f;
    }
}

static size_t enc_tail_smem(bool bwd) {
    return ((size_t)kHidden * kTailPitch + (size_t)kTailRows * kHidden * (bwd ? 3 : 2) + (bwd ? kTailRows * 2 : 0)) * sizeof(float);
}
constexpr int kTailStride = kHidden * kHidden + 2 * kHidden + kHidden + 4;      // floats per CTA partial

// workspace whose A^T (the encoder.l1 weight gradient's operand) the last encoder_forward of this thread transposed on a side
// stream; a backward pass on any other workspace, or after a forward that ran without side streams, transposes it itself
static thread_local const float* g_At_ws = nullptr;

static bool l1_wgrad_on_tc(int M) {
    static const bool tc_off = getenv("PAIG_NO_TCGEN05") != nullptr;
    static const int tc_mask = getenv("PAIG_TC_MASK") ? atoi(getenv("PAIG_TC_MASK")) : 6;
    return !tc_off && (tc_mask & 2) && M >= 128 && (M % 4) == 0;
}

int encoder_forward(const paig_task* t, const paig_params* p, const Layout& L, const float* x, long seq_stride,
                    int fps, float* enc_pos_out, float* enc_masks_out, float* masked_out, float* ws, cudaStream_t st) {
    const Dims& d = L.d;
    if (L.N <= 0) return 0;
    // The mask stage indexes frames as x + (f/fps)*seq_stride + (f%fps)*CHW; conv3x3 wants a plain batch stride,
    // so when the encoded frames are only a prefix of each sequence they are gathered once (kept for backward).
    const float* frames = x;
    static const bool layerwise0 = getenv("PAIG_UNET_LAYERWISE") != nullptr;
    const bool no_gather = (t->flags & PAIG_FLAG_INFERENCE) && !t->deep_unet && !layerwise0 && d.H <= 36;
    Side* sd = side_cur();
    const bool fused_path = !layerwise0 && !t->deep_unet;          // the fused forward reads x itself: only the
    if (seq_stride != (long)fps * d.CHW && !no_gather) {           // backward's c1 weight gradient wants the gathered copy
        float* dst = ws + L.frames;
        const long total4 = (long)L.N * d.CHW / 4;
        launch(gather_frames_kernel, dim3(cdiv(total4, 256)), dim3(256), 0, (sd && fused_path) ? sd->s1 : st, (const float4*)x,
               seq_stride / 4, fps, d.CHW / 4, total4, (float4*)dst);
        int rc0 = check_launch("gather_frames");
        if (rc0) return rc0;
        frames = dst;
    }
    // whole UNet in one persistent kernel when a frame's activations fit in shared memory (ShallowUNet); the
    // per-layer interpreter otherwise (64x64 UNet) or when PAIG_UNET_LAYERWISE=1 (A/B measurements)
    static const bool layerwise = getenv("PAIG_UNET_LAYERWISE") != nullptr;
    int rc = layerwise ? -1 : unet_tc_forward(t, p, L, x, seq_stride, fps, ws, st);        // tcgen05 (32-px frames)
    if (rc < 0 && !layerwise) rc = unet_fused_forward(t, p, L, x, seq_stride, fps, ws, st);
    if (rc < 0) rc = unet_forward(t, p, L, ws, frames, st);
    if (rc) return rc;
    float* masks = ws + L.masks;
    switch (d.n) {
        case 1: rc = masks_fwd<1>(L, t->deep_unet, ws, x, seq_stride, fps, masks, masked_out, st); break;
        case 2: rc = masks_fwd<2>(L, t->deep_unet, ws, x, seq_stride, fps, masks, masked_out, st); break;
        default: rc = masks_fwd<3>(L, t->deep_unet, ws, x, seq_stride, fps, masks, masked_out, st); break;
    }
    if (rc) return rc;
    if (enc_masks_out &&
        cudaMemcpyAsync(enc_masks_out, masks, (size_t)L.N * (d.n + 1) * d.HW * sizeof(float), cudaMemcpyDeviceToDevice,
                        st) != cudaSuccess)
        return check_launch("copy enc_masks");
    const int M = d.n * L.N;
    g_At_ws = nullptr;
    if (sd && l1_wgrad_on_tc(M) && !(t->flags & PAIG_FLAG_INFERENCE)) {       // (an inference workspace has no room for it)
        g_At_ws = ws;                                                           // the backward pass on THIS workspace may skip it
        // the backward's encoder.l1 weight gradient wants A^T: transpose it now, beside the MLP / rollout / decoder chain
        sd->fork1();
        float* At = ws + L.tc_scratch + (size_t)L.K * kHidden + (size_t)kHidden * M;
        if ((rc = transpose(ws + L.A, At, M, L.K, sd->s1))) return rc;
    }
    {   // encoder.l1: tcgen05 3xTF32 when the shape qualifies (gemm_tc.cu), else the CUDA-core GEMM
        // Which encoder.l1 GEMMs run on tcgen05 (1 forward | 2 weight gradient | 4 data gradient).  Default 7.  The
        // forward product uses the DRAINED kernel: with one TMEM accumulation chain per K split the tensor core's
        // round-toward-zero accumulate shrinks every pre-activation by ~2e-6, which the decoder's loss gradient
        // amplifies past the 1e-4 parity bar (measured in round 1: failed); chains of 4 + compensation pass.
        static const int tc_mask = getenv("PAIG_TC_MASK") ? atoi(getenv("PAIG_TC_MASK")) : 7;
        const int sp = !(tc_mask & 1) ? -1 : gemm_tc_partials_drained(ws + L.A, p->enc_l1.w, M, kHidden, L.K, ws + L.partials,
                                                                      L.partials_floats, "tc_l1_fwd", st);
        if (sp == 0) return 2;
        if (sp > 0) {
            GemmArgs g;
            g.C = ws + L.H1; g.ldc = kHidden; g.M = M; g.N = kHidden; g.bias = p->enc_l1.b; g.epi = EPI_RELU;
            g.splitk_ws = ws + L.partials;
            if ((rc = gemm_fold_partials(g, sp, st))) return rc;
        } else if ((rc = linear_forward(ws + L.A, p->enc_l1.w, p->enc_l1.b, ws + L.H1, M, L.K, kHidden, EPI_RELU, st,
                                        ws + L.partials, L.partials_floats, "sgemm_l1_fwd")))
            return rc;
    }
    static const bool tail_gemm = getenv("PAIG_TAIL_GEMM") != nullptr;
    if (!tail_gemm) {                                  // l2 + l3 + position head in one launch
        launch(enc_tail_fwd_kernel, dim3(cdiv(M, kTailRows)), dim3(256), enc_tail_smem(false), st, (const float*)(ws + L.H1), M,
               L.N, d.n, (float)d.H * 0.5f, (const float*)p->enc_l2.w, (const float*)p->enc_l2.b, (const float*)p->enc_l3.w,
               (const float*)p->enc_l3.b, ws + L.H2, ws + L.O3, ws + L.enc_pos, enc_pos_out);
        return check_launch("enc_tail_fwd");
    }
    if ((rc = linear_forward(ws + L.H1, p->enc_l2.w, p->enc_l2.b, ws + L.H2, M, kHidden, kHidden, EPI_RELU, st)))
        return rc;
    if ((rc = linear_forward(ws + L.H2, p->enc_l3.w, p->enc_l3.b, ws + L.O3, M, kHidden, 2, EPI_NONE, st))) return rc;
    launch(pos_head_kernel, dim3(cdiv(L.N * 2 * d.n, 256)), dim3(256), 0, st, (const float*)(ws + L.O3), L.N, d.n,
           (float)d.H * 0.5f, ws + L.enc_pos, enc_pos_out);
    return check_launch("pos_head");
}

// d_enc_pos: [N, 2n].  Writes every encoder gradient (UNet convs, l1..l3).
int encoder_backward(const paig_task* t, const paig_params* p, const paig_params* g, const Layout& L, const float* x,
                     long seq_stride, int fps, const float* d_enc_pos, float* ws, cudaStream_t st) {
    const Dims& d = L.d;
    if (L.N <= 0) return 0;
    int rc;
    const int M = d.n * L.N;
    static const bool tail_gemm = getenv("PAIG_TAIL_GEMM") != nullptr;
    const int tail_ctas = cdiv(M, kTailRows);
    if (!tail_gemm && (size_t)tail_ctas * kTailStride <= L.partials_floats) {
        // position head + l3 + l2 backward in one launch; per-CTA partials folded in fixed order
        float* part = ws + L.partials;
        launch(enc_tail_bwd_kernel, dim3(tail_ctas), dim3(256), enc_tail_smem(true), st, d_enc_pos, M, L.N, d.n,
               (float)d.H * 0.5f, (const float*)p->enc_l2.w, (const float*)p->enc_l3.w, (const float*)(ws + L.H1),
               (const float*)(ws + L.H2), (const float*)(ws + L.O3), ws + L.dH1, part, kTailStride);
        if ((rc = check_launch("enc_tail_bwd"))) return rc;
        ReduceBatch folds;
        const float* pW3 = part + kHidden * kHidden, *pb2 = pW3 + 2 * kHidden, *pb3 = pb2 + kHidden;
        if (g->enc_l2.w) folds.add(part, tail_ctas, kTailStride, kHidden * kHidden, g->enc_l2.w, 0, nullptr);
        if (g->enc_l3.w) folds.add(pW3, tail_ctas, kTailStride, 2 * kHidden, g->enc_l3.w, 0, nullptr);
        if (g->enc_l2.b) folds.add(pb2, tail_ctas, kTailStride, kHidden, g->enc_l2.b, 0, nullptr);
        if (g->enc_l3.b) folds.add(pb3, tail_ctas, kTailStride, 2, g->enc_l3.b, 0, nullptr);
        if ((rc = reduce_partials_batch(folds, st))) return rc;
    } else {
    launch(pos_head_bwd_kernel, dim3(cdiv(L.N * 2 * d.n, 256)), dim3(256), 0, st, (const float*)(ws + L.O3), L.N, d.n,
           (float)d.H * 0.5f, d_enc_pos, ws + L.dO3);
    if ((rc = check_launch("pos_head_bwd"))) return rc;
    // l3
    if ((rc = linear_wgrad(ws + L.dO3, ws + L.H2, g->enc_l3.w, g->enc_l3.b, M, kHidden, 2, st, ws + L.partials,
                           L.partials_floats)))
        return rc;
    if ((rc = linear_dgrad(ws + L.dO3, p->enc_l3.w, ws + L.dH2, M, kHidden, 2, EPI_MASK_RELU, ws + L.H2, st))) return rc;
    // l2
    if ((rc = linear_wgrad(ws + L.dH2, ws + L.H1, g->enc_l2.w, g->enc_l2.b, M, kHidden, kHidden, st, ws + L.partials,
                           L.partials_floats)))
        return rc;
    if ((rc = linear_dgrad(ws + L.dH2, p->enc_l2.w, ws + L.dH1, M, kHidden, kHidden, EPI_MASK_RELU, ws + L.H1, st)))
        return rc;
    }
    // l1: the weight gradient (transposes + tcgen05 GEMM + fold + bias column sums, ~110 us) does not feed the data
    // gradient: in the fused step it runs on a side stream beside the data-gradient GEMM and the UNet backward
    {
        Side* sd = side_cur();
        cudaStream_t wst = sd ? sd->s1 : st;
        float* Wt = ws + L.tc_scratch;                       // [K][200]
        float* dHt = Wt + (size_t)L.K * kHidden;             // [200][M]
        float* At = dHt + (size_t)kHidden * M;               // [K][M]   (already filled by encoder_forward when g_At_ws == ws)
        float* wpart = sd ? ws + L.partials2 : ws + L.partials;
        const size_t wpart_floats = sd ? L.partials2_floats : L.partials_floats;
        int sp = -1;
        static const bool tc_off = getenv("PAIG_NO_TCGEN05") != nullptr;
        static const int tc_mask = getenv("PAIG_TC_MASK") ? atoi(getenv("PAIG_TC_MASK")) : 6;
        if (sd) sd->fork1();
        if (l1_wgrad_on_tc(M)) {
            // dW1[200,K] = dH1^T[200,M] . A^T[K,M]^T : both operands transposed once so that M is the contiguous K axis
            if ((rc = transpose(ws + L.dH1, dHt, M, kHidden, wst))) return rc;
            if (!(sd && g_At_ws == ws) && (rc = transpose(ws + L.A, At, M, L.K, wst))) return rc;
            sp = gemm_tc_partials(dHt, At, kHidden, L.K, M, false, wpart, wpart_floats, "tc_l1_wgrad", wst);
            if (sp == 0) return 2;
        }
        if (sp > 0) {
            GemmArgs gg;
            gg.C = g->enc_l1.w; gg.ldc = L.K; gg.M = kHidden; gg.N = L.K; gg.splitk_ws = wpart;
            if ((rc = gemm_fold_partials(gg, sp, wst))) return rc;
            if (g->enc_l1.b && (rc = colsum(ws + L.dH1, M, kHidden, kHidden, g->enc_l1.b, wst))) return rc;
        } else if ((rc = linear_wgrad(ws + L.dH1, ws + L.A, g->enc_l1.w, g->enc_l1.b, M, L.K, kHidden, wst, wpart,
                                      wpart_floats, "sgemm_l1_wgrad")))
            return rc;
#ifndef PAIG_EMU
        // every gradient except the UNet conv layers' is final here (the VariableFromNetwork ones were enqueued on the
        // same side stream earlier): the data-parallel caller starts their all-reduce behind this event
        if (g_early_event) cudaEventRecord((cudaEvent_t)g_early_event, wst);
#endif
        // dA[M,K] = dH1[M,200] . (W1^T)[K,200]^T : one K split, written in place
        sp = -1;
        if (!tc_off && (tc_mask & 4) && M >= 128) {
            if ((rc = transpose(p->enc_l1.w, Wt, kHidden, L.K, st))) return rc;
            sp = gemm_tc_partials(ws + L.dH1, Wt, M, L.K, kHidden, true, ws + L.dA, (size_t)M * L.K, "tc_l1_dgrad", st);
            if (sp == 0) return 2;
            if (sp > 1) { set_error("tc_l1_dgrad: unexpected K split"); return 1; }
        }
        if (sp < 0 && (rc = linear_dgrad(ws + L.dH1, p->enc_l1.w, ws + L.dA, M, L.K, kHidden, EPI_NONE, nullptr, st, nullptr, 0,
                                         "sgemm_l1_dgrad")))
            return rc;
    }
    switch (d.n) {
        case 1: rc = masks_bwd<1>(L, t->deep_unet, ws, x, seq_stride, fps, st); break;
        case 2: rc = masks_bwd<2>(L, t->deep_unet, ws, x, seq_stride, fps, st); break;
        default: rc = masks_bwd<3>(L, t->deep_unet, ws, x, seq_stride, fps, st); break;
    }
    if (rc) return rc;
    const float* frames = seq_stride != (long)fps * d.CHW ? ws + L.frames : x;   // gathered by encoder_forward
    return unet_backward(t, p, g, L, ws, frames, st);
}

// ---------------------------------------------------------------------------------------------------------
// VelocityEncoder (blocks.py:31-49) and the initial rollout state (physics_models.py:220-228)
// ---------------------------------------------------------------------------------------------------------
// Vin[o*B + b][tau*2 + c] = enc_pos[b, tau, 2o+c]            (alt_vel: difference of consecutive steps)
__global__ void __launch_bounds__(256) vel_gather_kernel(const float* __restrict__ enc_pos, int B, int n, int e, int in,
                                                         int alt, float* __restrict__ vin) {
    const int cols = alt ? 2 * (in - 1) : 2 * in;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * B * cols) return;
    const int col = idx % cols, row = idx / cols, o = row / B, b = row % B, tau = col >> 1, c = col & 1;
    const float* ep = enc_pos + ((long)b * e) * 2 * n + 2 * o + c;
    vin[idx] = alt ? ep[(long)(tau + 1) * 2 * n] - ep[(long)tau * 2 * n] : ep[(long)tau * 2 * n];
}
// seq[b, 0, :] = [ enc_pos[b, in-1, :] | vel ]   with vel[b, 2o+c] = vout[o*B+b, c]  (zeros when in == 1)
__global__ void __launch_bounds__(256) state0_kernel(const float* __restrict__ enc_pos, const float* __restrict__ vout,
                                                     int B, int n, int e, int in, int steps, float* __restrict__ seq) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * 4 * n) return;
    const int b = idx / (4 * n), k = idx % (4 * n);
    float v;
    if (k < 2 * n) v = enc_pos[((long)b * e + (in - 1)) * 2 * n + k];
    else {
        const int kk = k - 2 * n;
        v = vout ? vout[((long)(kk >> 1) * B + b) * 2 + (kk & 1)] : 0.f;
    }
    seq[(long)b * (steps + 1) * 4 * n + k] = v;
}
__global__ void __launch_bounds__(256) state0_bwd_kernel(const float* __restrict__ d_state0, int B, int n, int e, int in,
                                                         float* __restrict__ d_enc_pos, float* __restrict__ dvout) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * 4 * n) return;
    const int b = idx / (4 * n), k = idx % (4 * n);
    const float v = d_state0[idx];
    if (k < 2 * n) d_enc_pos[((long)b * e + (in - 1)) * 2 * n + k] += v;
    else if (dvout) {
        const int kk = k - 2 * n;
        dvout[((long)(kk >> 1) * B + b) * 2 + (kk & 1)] = v;
    }
}
// adjoint of vel_gather: one thread per (b, o, c) walks tau, so no atomics
__global__ void __launch_bounds__(256) vel_scatter_kernel(const float* __restrict__ dvin, int B, int n, int e, int in,
                                                          int alt, float* __restrict__ d_enc_pos) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * 2 * n) return;
    const int b = idx / (2 * n), k = idx % (2 * n), o = k >> 1, c = k & 1;
    const int cols = alt ? 2 * (in - 1) : 2 * in;
    const float* dv = dvin + ((long)o * B + b) * cols + c;
    float* dp = d_enc_pos + ((long)b * e) * 2 * n + k;
    if (alt) {
        for (int tau = 0; tau < in - 1; ++tau) {
            const float v = dv[2 * tau];
            dp[(long)(tau + 1) * 2 * n] += v;
            dp[(long)tau * 2 * n] -= v;
        }
    } else {
        for (int tau = 0; tau < in; ++tau) dp[(long)tau * 2 * n] += dv[2 * tau];
    }
}

// ---- fused VelocityEncoder MLP (blocks.py:43-48: Linear(2 in,100) tanh Linear(100,100) tanh Linear(100,2)) -------------
// The problem is tiny (n B rows x 11 k MACs) and was 5 + 14 launch-latency-bound GEMM / gather / scatter launches; two
// kernels do it: forward (8 rows per CTA, weights transposed into shared memory so unit j of a layer is lane j) also
// gathers the inputs and writes the rollout's initial state; backward is ONE CTA that walks the rows in chunks of 8 and
// keeps every weight-gradient entry in a register of its owning thread (fixed order => deterministic), then scatters the
// input gradient into d enc_pos.
constexpr int kVelRows = 8;
constexpr int kVelMaxIn = 16;             // 2 * input_steps

__global__ void __launch_bounds__(128) vel_mlp_fwd_kernel(const float* __restrict__ enc_pos, int B, int n, int e, int in,
                                                          int steps, const float* __restrict__ W1, const float* __restrict__ b1,
                                                          const float* __restrict__ W2, const float* __restrict__ b2,
                                                          const float* __restrict__ W3, const float* __restrict__ b3,
                                                          float* __restrict__ vin, float* __restrict__ v1, float* __restrict__ v2,
                                                          float* __restrict__ vout, float* __restrict__ seq) {
    constexpr int HID = kVelHidden;
    PAIG_DYN_SMEM(float, smem);
    float* sW2t = smem;                               // [k][j]
    float* sW1t = sW2t + HID * HID;                   // [k][j]
    float (*sIn)[kVelMaxIn] = reinterpret_cast<float (*)[kVelMaxIn]>(sW1t + kVelMaxIn * HID);
    float (*sH1)[HID] = reinterpret_cast<float (*)[HID]>(&sIn[kVelRows][0]);
    float (*sH2)[HID] = reinterpret_cast<float (*)[HID]>(&sH1[kVelRows][0]);
    const int tid = threadIdx.x, cols = 2 * in, M = n * B;
    const int r0 = blockIdx.x * kVelRows;
    // (loads batched 8 deep: one global load per trip exposed the L2 latency 78 times in a row -- 22 us for a 2 MFLOP kernel)
#pragma unroll 8
    for (int i = tid; i < HID * HID; i += 128) sW2t[(i % HID) * HID + i / HID] = W2[i];
#pragma unroll 4
    for (int i = tid; i < HID * cols; i += 128) sW1t[(i % cols) * HID + i / cols] = W1[i];
    for (int i = tid; i < kVelRows * cols; i += 128) {
        const int r = i / cols, col = i % cols, row = r0 + r;
        float v = 0.f;
        if (row < M) {
            const int o = row / B, b = row % B, tau = col >> 1, c = col & 1;
            v = enc_pos[((long)b * e + tau) * 2 * n + 2 * o + c];
            vin[(long)row * cols + col] = v;
        }
        sIn[r][col] = v;
    }
    __syncthreads();
    if (tid < HID) {
        float acc[kVelRows];
#pragma unroll
        for (int r = 0; r < kVelRows; ++r) acc[r] = 0.f;
        for (int k = 0; k < cols; ++k) {
            const float w = sW1t[k * HID + tid];
#pragma unroll
            for (int r = 0; r < kVelRows; ++r) acc[r] += sIn[r][k] * w;
        }
        const float bb = b1[tid];
#pragma unroll
        for (int r = 0; r < kVelRows; ++r) {
            const float h = tanhf(acc[r] + bb);
            sH1[r][tid] = h;
            if (r0 + r < M) v1[(long)(r0 + r) * HID + tid] = h;
        }
    }
    __syncthreads();
    if (tid < HID) {
        float acc[kVelRows];
#pragma unroll
        for (int r = 0; r < kVelRows; ++r) acc[r] = 0.f;
        for (int k = 0; k < HID; ++k) {
            const float w = sW2t[k * HID + tid];
#pragma unroll
            for (int r = 0; r < kVelRows; ++r) acc[r] += sH1[r][k] * w;
        }
        const float bb = b2[tid];
#pragma unroll
        for (int r = 0; r < kVelRows; ++r) {
            const float h = tanhf(acc[r] + bb);
            sH2[r][tid] = h;
            if (r0 + r < M) v2[(long)(r0 + r) * HID + tid] = h;
        }
    }
    __syncthreads();
    if (tid < kVelRows * 2) {
        const int r = tid >> 1, c = tid & 1, row = r0 + r;
        if (row < M) {
            float s = 0.f;
            for (int k = 0; k < HID; ++k) s += sH2[r][k] * W3[c * HID + k];
            s += b3[c];
            vout[(long)row * 2 + c] = s;
            const int o = row / B, b = row % B;
            float* sq = seq + (long)b * (steps + 1) * 4 * n;           // physics_models.py:225-228: [pos(in-1) | vel]
            sq[2 * o + c] = enc_pos[((long)b * e + (in - 1)) * 2 * n + 2 * o + c];
            sq[2 * n + 2 * o + c] = s;
        }
    }
}

__global__ void __launch_bounds__(512) vel_mlp_bwd_kernel(const float* __restrict__ d_state0, int B, int n, int e, int in,
                                                          const float* __restrict__ W1, const float* __restrict__ W2,
                                                          const float* __restrict__ W3, const float* __restrict__ vin,
                                                          const float* __restrict__ v1, const float* __restrict__ v2,
                                                          float* __restrict__ partials, int stride,
                                                          float* __restrict__ d_enc_pos) {
    constexpr int HID = kVelHidden, NT = 512, R = kVelRows;
    constexpr int W2_PER = (HID * HID + NT - 1) / NT;                  // 20 entries of dW2 per thread
    PAIG_DYN_SMEM(float, smem);
    float* sW2 = smem;                                                 // [j][k]
    float* sW1 = sW2 + HID * HID;
    float* sW3 = sW1 + HID * kVelMaxIn;
    float (*sH1)[HID] = reinterpret_cast<float (*)[HID]>(sW3 + 2 * HID);
    float (*sH2)[HID] = reinterpret_cast<float (*)[HID]>(&sH1[R][0]);
    float (*sDz2)[HID] = reinterpret_cast<float (*)[HID]>(&sH2[R][0]);
    float (*sDz1)[HID] = reinterpret_cast<float (*)[HID]>(&sDz2[R][0]);
    float (*sIn)[kVelMaxIn] = reinterpret_cast<float (*)[kVelMaxIn]>(&sDz1[R][0]);
    float (*sDz3)[2] = reinterpret_cast<float (*)[2]>(&sIn[R][0]);
    const int tid = threadIdx.x, cols = 2 * in, M = n * B;
#pragma unroll 8
    for (int i = tid; i < HID * HID; i += NT) sW2[i] = W2[i];
#pragma unroll 4
    for (int i = tid; i < HID * cols; i += NT) sW1[i] = W1[i];
    for (int i = tid; i < 2 * HID; i += NT) sW3[i] = W3[i];
    float aW2[W2_PER], aW1[4], aW3 = 0.f, ab = 0.f;                    // gradients owned by this thread
#pragma unroll
    for (int i = 0; i < W2_PER; ++i) aW2[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) aW1[i] = 0.f;
    __syncthreads();
    {
        const int r0 = blockIdx.x * R;                                 // one chunk of rows per CTA
        const int nr = min(R, M - r0);
        for (int i = tid; i < R * HID; i += NT) {
            const int r = i / HID, k = i % HID;
            sH1[r][k] = r < nr ? v1[(long)(r0 + r) * HID + k] : 0.f;
            sH2[r][k] = r < nr ? v2[(long)(r0 + r) * HID + k] : 0.f;
        }
        for (int i = tid; i < R * cols; i += NT) sIn[i / cols][i % cols] = i / cols < nr ? vin[(long)(r0 + i / cols) * cols + i % cols] : 0.f;
        if (tid < R * 2) {
            const int r = tid >> 1, c = tid & 1, row = r0 + r;
            float v = 0.f;
            if (r < nr) { const int o = row / B, b = row % B; v = d_state0[(long)b * 4 * n + 2 * n + 2 * o + c]; }
            sDz3[r][c] = v;
        }
        __syncthreads();
        // layer 3: dW3[c][k] (thread c*100+k), db3[c] (threads 200, 201); dz2 = (W3^T dz3) (1 - h2^2)
        if (tid < 2 * HID) {
            const int c = tid / HID, k = tid % HID;
#pragma unroll
            for (int r = 0; r < R; ++r) aW3 += sDz3[r][c] * sH2[r][k];
        } else if (tid < 2 * HID + 2) {
#pragma unroll
            for (int r = 0; r < R; ++r) ab += sDz3[r][tid - 2 * HID];
        }
        for (int i = tid; i < R * HID; i += NT) {
            const int r = i / HID, k = i % HID;
            const float h = sH2[r][k];
            sDz2[r][k] = (sW3[k] * sDz3[r][0] + sW3[HID + k] * sDz3[r][1]) * (1.f - h * h);
        }
        __syncthreads();
        // layer 2: dW2 entries e = tid + 512 i -> (j, k); db2[j] (threads 256..355); dz1 = (W2^T dz2) (1 - h1^2)
#pragma unroll
        for (int i = 0; i < W2_PER; ++i) {
            const int en = tid + NT * i;
            if (en < HID * HID) {
                const int j = en / HID, k = en % HID;
                float s = aW2[i];
#pragma unroll
                for (int r = 0; r < R; ++r) s += sDz2[r][j] * sH1[r][k];
                aW2[i] = s;
            }
        }
        if (tid >= 256 && tid < 256 + HID) {
#pragma unroll
            for (int r = 0; r < R; ++r) ab += sDz2[r][tid - 256];
        }
        for (int i = tid; i < R * HID; i += NT) {
            const int r = i / HID, k = i % HID;
            float s = 0.f;
            for (int j = 0; j < HID; ++j) s += sW2[j * HID + k] * sDz2[r][j];
            const float h = sH1[r][k];
            sDz1[r][k] = s * (1.f - h * h);
        }
        __syncthreads();
        // layer 1: dW1 entries e = tid + 512 i < 100 cols; db1[j] (threads 384..483); d vin scattered into d enc_pos
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int en = tid + NT * i;
            if (en < HID * cols) {
                const int j = en / cols, k = en % cols;
                float s = aW1[i];
#pragma unroll
                for (int r = 0; r < R; ++r) s += sDz1[r][j] * sIn[r][k];
                aW1[i] = s;
            }
        }
        if (tid >= 384 && tid < 384 + HID) {
#pragma unroll
            for (int r = 0; r < R; ++r) ab += sDz1[r][tid - 384];
        }
        if (tid < R * cols) {
            const int r = tid / cols, col = tid % cols, row = r0 + r;
            if (r < nr) {
                float s = 0.f;
                for (int j = 0; j < HID; ++j) s += sW1[j * cols + col] * sDz1[r][j];
                const int o = row / B, b = row % B, tau = col >> 1, c = col & 1;
                // the position half of d state0 goes to d enc_pos[b, in-1, :] (physics_models.py:225): same owner
                if (tau == in - 1) s += d_state0[(long)b * 4 * n + 2 * o + c];
                d_enc_pos[((long)b * e + tau) * 2 * n + 2 * o + c] += s;     // each (b, tau, o, c) has one owner
            }
        }
        __syncthreads();
    }
    // this CTA's partial sums: [dW2 | dW1 | dW3 | db1 | db2 | db3], folded in fixed order by reduce_partials_batch
    float* out = partials + (size_t)blockIdx.x * stride;
    float* oW1 = out + HID * HID, *oW3 = oW1 + HID * cols, *ob1 = oW3 + 2 * HID, *ob2 = ob1 + HID, *ob3 = ob2 + HID;
    if (tid < 2 * HID) oW3[tid] = aW3;
    else if (tid < 2 * HID + 2) ob3[tid - 2 * HID] = ab;
    if (tid >= 256 && tid < 256 + HID) ob2[tid - 256] = ab;
    if (tid >= 384 && tid < 384 + HID) ob1[tid - 384] = ab;
#pragma unroll
    for (int i = 0; i < W2_PER; ++i)
        if (tid + NT * i < HID * HID) out[tid + NT * i] = aW2[i];
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (tid + NT * i < HID * cols) oW1[tid + NT * i] = aW1[i];
}

int velocity_forward(const paig_task* t, const paig_params* p, const Layout& L, const float* enc_pos, float* ws,
                     cudaStream_t st) {
    const Dims& d = L.d;
    if (L.B <= 0) return 0;
    int rc = 0;
    const float* vout = nullptr;
    if (d.in > 1) {                                                   // physics_models.py:220-223
        const int M = d.n * L.B, cols = t->alt_vel ? 2 * (d.in - 1) : 2 * d.in;
        static const bool vel_gemm = getenv("PAIG_VEL_GEMM") != nullptr;
        if (!t->alt_vel && !vel_gemm && cols <= kVelMaxIn) {            // one fused launch (gather + MLP + initial state)
            const size_t smem = ((size_t)kVelHidden * kVelHidden + (size_t)kVelMaxIn * kVelHidden +
                                 (size_t)kVelRows * (kVelMaxIn + 2 * kVelHidden)) * sizeof(float);
            launch(vel_mlp_fwd_kernel, dim3(cdiv(M, kVelRows)), dim3(128), smem, st, enc_pos, L.B, d.n, d.e, d.in, d.steps,
                   (const float*)p->vel[0].w, (const float*)p->vel[0].b, (const float*)p->vel[1].w, (const float*)p->vel[1].b,
                   (const float*)p->vel[2].w, (const float*)p->vel[2].b, ws + L.vin, ws + L.v1, ws + L.v2, ws + L.vout,
                   ws + L.seq);
            return check_launch("vel_mlp_fwd");
        }
        launch(vel_gather_kernel, dim3(cdiv(M * cols, 256)), dim3(256), 0, st, enc_pos, L.B, d.n, d.e, d.in,
               (int)t->alt_vel, ws + L.vin);
        if ((rc = check_launch("vel_gather"))) return rc;
        if (t->alt_vel) {
            rc = linear_forward(ws + L.vin, p->vel[0].w, p->vel[0].b, ws + L.vout, M, cols, 2, EPI_NONE, st);
        } else {
            rc = linear_forward(ws + L.vin, p->vel[0].w, p->vel[0].b, ws + L.v1, M, cols, kVelHidden, EPI_TANH, st);
            if (!rc) rc = linear_forward(ws + L.v1, p->vel[1].w, p->vel[1].b, ws + L.v2, M, kVelHidden, kVelHidden, EPI_TANH, st);
            if (!rc) rc = linear_forward(ws + L.v2, p->vel[2].w, p->vel[2].b, ws + L.vout, M, kVelHidden, 2, EPI_NONE, st);
        }
        if (rc) return rc;
        vout = ws + L.vout;
    }
    launch(state0_kernel, dim3(cdiv(L.B * 4 * d.n, 256)), dim3(256), 0, st, enc_pos, vout, L.B, d.n, d.e, d.in, d.steps,
           ws + L.seq);
    return check_launch("state0");
}

// d_state0 [B,4n] -> accumulates into d_enc_pos [B,e,2n]; writes the velocity-encoder gradients.
int velocity_backward(const paig_task* t, const paig_params* p, const paig_params* g, const Layout& L,
                      const float* d_state0, float* d_enc_pos, float* ws, cudaStream_t st) {
    const Dims& d = L.d;
    if (L.B <= 0) return 0;
    int rc;
    const bool has_vel = d.in > 1;
    {
        static const bool vel_gemm = getenv("PAIG_VEL_GEMM") != nullptr;
        const int M = d.n * L.B, cols = 2 * d.in;
        const int HID = kVelHidden;
        const int stride = HID * HID + HID * cols + 2 * HID + 2 * HID + 2 + 2;          // floats per CTA partial (padded)
        const int ctas = cdiv(M, kVelRows);
        if (has_vel && !t->alt_vel && !vel_gemm && cols <= kVelMaxIn && (size_t)ctas * stride <= L.partials_floats) {
            const size_t smem = ((size_t)HID * HID + (size_t)HID * kVelMaxIn + 2 * HID +
                                 (size_t)kVelRows * (4 * HID + kVelMaxIn + 2)) * sizeof(float);
            float* part = ws + L.partials;
            launch(vel_mlp_bwd_kernel, dim3(ctas), dim3(512), smem, st, d_state0, L.B, d.n, d.e, d.in, (const float*)p->vel[0].w,
                   (const float*)p->vel[1].w, (const float*)p->vel[2].w, (const float*)(ws + L.vin), (const float*)(ws + L.v1),
                   (const float*)(ws + L.v2), part, stride, d_enc_pos);
            if ((rc = check_launch("vel_mlp_bwd"))) return rc;
            ReduceBatch folds;                                        // (NULL destinations: that gradient is not wanted)
            const float* pW1 = part + HID * HID, *pW3 = pW1 + HID * cols, *pb1 = pW3 + 2 * HID, *pb2 = pb1 + HID, *pb3 = pb2 + HID;
            if (g->vel[1].w) folds.add(part, ctas, stride, HID * HID, g->vel[1].w, 0, nullptr);
            if (g->vel[0].w) folds.add(pW1, ctas, stride, HID * cols, g->vel[0].w, 0, nullptr);
            if (g->vel[2].w) folds.add(pW3, ctas, stride, 2 * HID, g->vel[2].w, 0, nullptr);
            if (g->vel[0].b) folds.add(pb1, ctas, stride, HID, g->vel[0].b, 0, nullptr);
            if (g->vel[1].b) folds.add(pb2, ctas, stride, HID, g->vel[1].b, 0, nullptr);
            if (g->vel[2].b) folds.add(pb3, ctas, stride, 2, g->vel[2].b, 0, nullptr);
            return reduce_partials_batch(folds, st);
        }
    }
    launch(state0_bwd_kernel, dim3(cdiv(L.B * 4 * d.n, 256)), dim3(256), 0, st, d_state0, L.B, d.n, d.e, d.in, d_enc_pos,
           has_vel ? ws + L.dvout : (float*)nullptr);
    if ((rc = check_launch("state0_bwd"))) return rc;
    if (!has_vel) return 0;
    const int M = d.n * L.B, cols = t->alt_vel ? 2 * (d.in - 1) : 2 * d.in;
    if (t->alt_vel) {
        if ((rc = linear_wgrad(ws + L.dvout, ws + L.vin, g->vel[0].w, g->vel[0].b, M, cols, 2, st, ws + L.partials, L.partials_floats))) return rc;
        if ((rc = linear_dgrad(ws + L.dvout, p->vel[0].w, ws + L.dvin, M, cols, 2, EPI_NONE, nullptr, st))) return rc;
    } else {
        if ((rc = linear_wgrad(ws + L.dvout, ws + L.v2, g->vel[2].w, g->vel[2].b, M, kVelHidden, 2, st, ws + L.partials, L.partials_floats))) return rc;
        if ((rc = linear_dgrad(ws + L.dvout, p->vel[2].w, ws + L.dv2, M, kVelHidden, 2, EPI_MASK_TANH, ws + L.v2, st)))
            return rc;
        if ((rc = linear_wgrad(ws + L.dv2, ws + L.v1, g->vel[1].w, g->vel[1].b, M, kVelHidden, kVelHidden, st, ws + L.partials, L.partials_floats))) return rc;
        if ((rc = linear_dgrad(ws + L.dv2, p->vel[1].w, ws + L.dv1, M, kVelHidden, kVelHidden, EPI_MASK_TANH, ws + L.v1,
                               st)))
            return rc;
        if ((rc = linear_wgrad(ws + L.dv1, ws + L.vin, g->vel[0].w, g->vel[0].b, M, cols, kVelHidden, st, ws + L.partials, L.partials_floats))) return rc;
        if ((rc = linear_dgrad(ws + L.dv1, p->vel[0].w, ws + L.dvin, M, cols, kVelHidden, EPI_NONE, nullptr, st)))
            return rc;
    }
    launch(vel_scatter_kernel, dim3(cdiv(L.B * 2 * d.n, 256)), dim3(256), 0, st, (const float*)(ws + L.dvin), L.B, d.n,
           d.e, d.in, (int)t->alt_vel, d_enc_pos);
    return check_launch("vel_scatter");
}

}  // namespace paig
