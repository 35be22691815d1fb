/* ldpc_oracle.c -- plain-C restatement of the reference's hot path (see ldpc_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY: the checker, never the thing shipped or measured (except as the reported
 * CPU baseline of bench.py).  Scalar code, 32 lanes (= frames) per group exactly like the reference's
 * byte lanes of one __m256i; rows are walked in PosNoeudsVariable order (block row 0..11, 256 rows each),
 * which is the reference's schedule.  Each block cites the reference lines it restates.
 */
#define _GNU_SOURCE
#include "ldpc_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "ldpc_code_tables.h"

#define N LDPC_N
#define M LDPC_M
#define K LDPC_K
#define QZ LDPC_Z
#define E LDPC_E
#define LANES 32

/* Constants_SSE.h:20-25: 6-bit APP, 4-bit messages */
#define SAT_POS_VAR 31
#define SAT_NEG_VAR (-31)
#define SAT_POS_MSG 7

static inline int sat8(int x) { return x > 127 ? 127 : (x < -128 ? -128 : x); }
static inline int usat8(int x) { return x > 255 ? 255 : (x < 0 ? 0 : x); }
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int iabs(int a) { return a < 0 ? -a : a; }

/* ------------------------------------------------------------------------------------------------ */
/* Shipped constants                                                                                 */
/* ------------------------------------------------------------------------------------------------ */

/* CDecoder_FAID.cpp:12-49 (FAID3), :51-88 (FAID32), :90-127 (FAID2); CDecoder_FAID_2B1C.cpp:12-47 (hybrid).
 * All four weight-class rows of every shipped table are equal. */
static const int8_t LUT_SETS[4][6][8] = {
    /* FAID3 */
    {{0, 1, 1, 2, 3, 3, 3, 3}, {0, 1, 1, 2, 3, 3, 3, 3}, {0, 1, 1, 2, 4, 4, 4, 4},
     {0, 1, 1, 3, 3, 4, 4, 4}, {0, 1, 1, 3, 3, 3, 6, 6}, {0, 1, 1, 3, 3, 3, 7, 7}},
    /* FAID32 */
    {{0, 1, 1, 2, 3, 3, 3, 3}, {0, 1, 1, 2, 3, 3, 3, 3}, {0, 1, 1, 2, 4, 4, 4, 4},
     {1, 1, 1, 1, 4, 4, 4, 4}, {1, 1, 1, 1, 5, 5, 5, 5}, {1, 1, 1, 1, 6, 6, 6, 6}},
    /* FAID2 */
    {{0, 0, 2, 2, 2, 2, 2, 2}, {0, 0, 2, 2, 2, 2, 2, 2}, {1, 1, 1, 3, 3, 3, 3, 3},
     {1, 1, 1, 4, 4, 4, 4, 4}, {1, 1, 1, 5, 5, 5, 5, 5}, {1, 1, 1, 6, 6, 6, 6, 6}},
    /* hybrid */
    {{0, 0, 1, 2, 3, 3, 3, 3}, {0, 1, 1, 2, 3, 3, 3, 3}, {0, 1, 1, 2, 3, 3, 3, 3},
     {0, 1, 1, 3, 3, 4, 4, 4}, {0, 1, 1, 3, 3, 3, 6, 6}, {0, 1, 1, 3, 3, 3, 7, 7}},
};
/* CDecoder_FAID.cpp:130-165, CDecoder_FAID_2B1C.cpp:50-85 */
static const int8_t LUT_EF[8] = {2, 3, 3, 4, 5, 6, 6, 7};

void ldpc_oracle_default_config(ldpc_b200_config* c, int method, int lut_variant) {
    memset(c, 0, sizeof *c);
    c->struct_size = sizeof *c;
    c->abi_version = LDPC_B200_ABI_VERSION;
    /* Profile.txt as shipped */
    c->snr_start = 3.0f; c->snr_pass = 0.1f; c->snr_end = 5.0f;
    c->decode_method = method;
    c->max_iteration = 6;
    c->mod_type = 2;
    c->interleave_mod_type = 1;
    c->factor_1 = 1; c->factor_2 = 6;
    if (method == 0) { c->factor_1 = 26; c->factor_2 = 26; } /* README.md:22, BASELINE config 1 */
    c->nb_frames = 32;
    c->scale = (method == 5) ? 12.5f : 13.0f;
    c->Z = 256;
    if (lut_variant < 0 || lut_variant > 3) lut_variant = (method == 5) ? 3 : 0;
    for (int it = 0; it < 6; ++it)
        for (int w = 0; w < 4; ++w)
            for (int a = 0; a < 8; ++a) {
                c->v2c_lut[it][w][a] = LUT_SETS[lut_variant][it][a];
                c->v2c_lut_ef[it][w][a] = LUT_EF[a];
            }
    /* CDecoder_FAID.cpp:5,192-195 (EF 0: count 0, thresh -1); CDecoder_FAID_2B1C.cpp:6,116-119 (EF 1: 50, 6) */
    c->ef_elimination = (method == 5) ? 1 : 0;
    c->ef_floor_err_count = (method == 5) ? 50 : 0;
    c->ef_floor_iter_thresh = (method == 5) ? 6 : -1;
    c->oms_floor_err_count = 100; /* CDecoder_OMS.cpp:26 */
    c->oms_floor_iter_thresh = 4; /* CDecoder_OMS.cpp:27 */
    c->oms_mode = 1;              /* CDecoder_OMS.cpp:3 */
    c->oms_offset = 1;            /* CDecoder_OMS.cpp:6 */
    c->regular_col_weight = 3;    /* CTool.h:6 */
    c->hard2_threshold = 13;      /* CDecoder_FAID_2B1C.cpp:6130 */
    c->puncture_tail = 384;       /* CLDPC.cpp:270-272 */
    c->code_rate = 0.8444444;     /* CLDPC.cpp:4780 */
    c->dtbf_delta = 1; c->dtbf_alpha = 1;
    switch (method) {
    case 2: c->bf_mode = LDPC_B200_BF_DTBF; c->bf_max_iter = 10; c->dtbf_L0 = 50; c->dtbf_L1 = 0; break;  /* CDecoder_FAID.cpp:167-170,208 */
    case 3: c->bf_mode = LDPC_B200_BF_PLAIN; c->bf_max_iter = 50; break;                                     /* CDecoder_OMSBF.cpp:30 */
    case 4: c->bf_mode = LDPC_B200_BF_DTBF; c->bf_max_iter = 50; c->dtbf_L0 = 0; c->dtbf_L1 = 50; break;    /* CDecoder_OMS_DTBF.cpp:6-9,35 */
    case 5: c->bf_mode = LDPC_B200_BF_2B1C; c->bf_max_iter = 10; c->dtbf_L0 = 100; c->dtbf_L1 = 0; break;   /* CDecoder_FAID_2B1C.cpp:87-90,128 */
    default: c->bf_mode = LDPC_B200_BF_NONE; c->bf_max_iter = 0; break;
    }
    c->device = 0; c->n_streams = 6; c->chunk_groups = 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* Code structure                                                                                    */
/* ------------------------------------------------------------------------------------------------ */

/* VN index of edge j of row `row` (PosNoeudsVariable order, Constants_SSE.h:29-3102) */
static inline int row_vn(int row, int j) {
    int b = row / QZ, r = row % QZ;
    int e = ldpc_layer_start[b] + j;
    return ldpc_circ_col[e] * QZ + ((ldpc_circ_shift[e] + r) % QZ);
}
static inline int row_deg(int row) { return ldpc_layer_deg[row / QZ]; }
/* column weight of a code bit (CLDPC.cpp:4998-5003, intended semantics) */
static inline int vn_weight(int n) { return ldpc_col_weight[n / QZ]; }
/* CDecoder_FAID.cpp:692-705 */
static inline int wclass(int weight) { return weight == 3 ? 0 : weight == 6 ? 1 : weight == 11 ? 2 : 3; }

/* ------------------------------------------------------------------------------------------------ */
/* Decoders                                                                                          */
/* ------------------------------------------------------------------------------------------------ */

typedef struct {
    int8_t L[N][LANES];    /* var_nodes: APP, lane = frame */
    int8_t msg[E][LANES];  /* var_msgs: C2V in edge order */
    uint8_t chk[M][LANES]; /* l_checksum_: row unsatisfied at iteration start */
    int8_t hard[N][LANES], hard2[N][LANES], hch[N][LANES], rec[N][LANES];
    uint8_t fv[N][LANES];
    uint8_t vote[N][LANES], era[N][LANES]; /* EF_ELIMINATION 2: flip_vote at iteration start, erased-this-iteration flag */
} work_t;

/* syndrome of hard decisions (bit = L > 0): fills chk, per-lane error_sum with the given saturation.
 * CDecoder_OMS.cpp:102-136 (unsigned saturation, VECTOR_ADDU_MASK), CDecoder_FAID.cpp:294-343 (signed, VECTOR_ADD_MASK) */
static void syndrome_llr(work_t* w, int err[LANES], int sat_max) {
    for (int f = 0; f < LANES; ++f) err[f] = 0;
    memset(w->vote, 0, sizeof w->vote); /* CDecoder_FAID.cpp:287-290: flip_vote is rebuilt in every iteration */
    memset(w->era, 0, sizeof w->era);   /* :623-628 */
    for (int row = 0; row < M; ++row) {
        int d = row_deg(row);
        for (int f = 0; f < LANES; ++f) {
            int x = 0;
            for (int j = 0; j < d; ++j) x ^= (w->L[row_vn(row, j)][f] > 0);
            w->chk[row][f] = (uint8_t)x;
            if (x) {
                err[f] = imin(err[f] + 1, sat_max);
                for (int j = 0; j < d; ++j) w->vote[row_vn(row, j)][f]++; /* :306-309 (VECTOR_ADDU_MASK; at most 12) */
            }
        }
    }
}

/* One row of the NMS decoder, CLDPC.cpp:293-406 */
static void row_nms(work_t* w, const ldpc_b200_config* c, int row, int ebase) {
    int d = row_deg(row);
    int odd = d & 1;
    for (int f = 0; f < LANES; ++f) {
        int v[LDPC_MAXDEG];
        int sign = 0, min1 = SAT_POS_VAR, min2 = SAT_POS_VAR; /* :296-297 */
        for (int j = 0; j < d; ++j) {
            int n = row_vn(row, j);
            int x = imax(sat8(w->L[n][f] - w->msg[ebase + j][f]), SAT_NEG_VAR); /* :330, lower clamp only */
            sign ^= (x < 0);                                                   /* :331-332 */
            int a = iabs(x);                                                   /* :333 */
            v[j] = x;
            int t = min1;
            min1 = imin(a, min1);            /* :336 */
            min2 = imin(min2, imax(t, a));   /* :337, CLDPC.h:68 */
        }
        /* :342-363: (min * factor) >> 5 on zero-extended u16 lanes, signed-saturating pack, then min with 7 */
        int c2 = (int)(((unsigned)(min1 & 0xff) * (unsigned)(c->factor_1 & 0xffff)) & 0xffff) >> 5;
        int c1 = (int)(((unsigned)(min2 & 0xff) * (unsigned)(c->factor_2 & 0xffff)) & 0xffff) >> 5;
        c2 = imin(imin(c2, 127), SAT_POS_MSG);
        c1 = imin(imin(c1, 127), SAT_POS_MSG);
        sign ^= odd; /* :374-378: 0xC0 for odd degree flips the sign bit, 0x40 for even does not */
        for (int j = 0; j < d; ++j) {
            int n = row_vn(row, j);
            int a = iabs(v[j]);
            int mag = (a == min1) ? c1 : c2;          /* :384-387 */
            int neg = sign ^ (v[j] < 0);               /* :388 */
            int m = neg ? -mag : mag;                  /* :389 */
            int l = imin(imax(sat8(v[j] + m), SAT_NEG_VAR), SAT_POS_VAR); /* :390-391 */
            w->msg[ebase + j][f] = (int8_t)m;
            w->L[n][f] = (int8_t)l;
        }
    }
}

/* selective offset of one minimum, CDecoder_OMS.cpp:386-432 (OMS_MODE 1) */
static int oms_offset(int m, int sel_active, int sel, int F1, int F2) {
    if (sel_active) {
        if (sel) {
            if (m < F2) m = sat8(m + 1);  /* msk_lt6 */
            if (m <= F1) m = sat8(m + 1); /* msk_le1 (tests the updated value) */
        } else {
            if (m > F1) m = sat8(m - 1);  /* msk_gt1 */
            if (m >= F2) m = sat8(m - 1); /* msk_ge6 */
        }
    } else {
        if (m > F1) m = sat8(m - 1);
        if (m >= F2) m = sat8(m - 1);
    }
    return m;
}

/* One row of the OMS decoder, CDecoder_OMS.cpp:334-478 (identical in CDecoder_OMSBF.cpp / CDecoder_OMS_DTBF.cpp) */
static void row_oms(work_t* w, const ldpc_b200_config* c, int row, int ebase, int remaining, const uint8_t lt_floor[LANES]) {
    int d = row_deg(row);
    int odd = d & 1;
    int F1 = (int8_t)c->factor_1, F2 = (int8_t)c->factor_2; /* VECTOR_SET1 on int8 lanes */
    for (int f = 0; f < LANES; ++f) {
        int v[LDPC_MAXDEG];
        int sign = 0, min1 = SAT_POS_VAR, min2 = SAT_POS_VAR;
        for (int j = 0; j < d; ++j) {
            int n = row_vn(row, j);
            int x = imax(sat8(w->L[n][f] - w->msg[ebase + j][f]), SAT_NEG_VAR); /* :371 */
            sign ^= (x < 0);
            int a = imin(iabs(x), SAT_POS_MSG); /* :374: clipped BEFORE the min search */
            v[j] = x;
            min2 = imin(min2, imax(min1, a));   /* :376 */
            min1 = imin(a, min1);               /* :377 */
        }
        int active = remaining <= c->oms_floor_iter_thresh; /* :389 */
        int sel = w->chk[row][f] && lt_floor[f];
        int m1, m2;
        if (c->oms_mode == 0) { /* OMS_MODE 0, CDecoder_OMS.cpp:383-385: may go negative for min < offset */
            m1 = sat8(min1 - (int8_t)c->oms_offset);
            m2 = sat8(min2 - (int8_t)c->oms_offset);
        } else {
            m1 = oms_offset(min1, active, sel, F1, F2);
            m2 = oms_offset(min2, active, sel, F1, F2);
        }
        int c1 = imin(m2, SAT_POS_MSG); /* :431 */
        int c2 = imin(m1, SAT_POS_MSG); /* :432 */
        sign ^= odd;
        for (int j = 0; j < d; ++j) {
            int n = row_vn(row, j);
            int a = iabs(v[j]);                  /* :453: UNCLIPPED |v| compared with the clipped min1 */
            int mag = (a == min1) ? c1 : c2;
            int neg = sign ^ (v[j] < 0);
            int m = neg ? -mag : mag;
            int l = imin(imax(sat8(v[j] + m), SAT_NEG_VAR), SAT_POS_VAR);
            w->msg[ebase + j][f] = (int8_t)m;
            w->L[n][f] = (int8_t)l;
        }
    }
}

/* One row of the LNS-FAID decoder, CDecoder_FAID.cpp:631-936 / CDecoder_FAID_2B1C.cpp:444-745 */
static void row_faid(work_t* w, const ldpc_b200_config* c, int row, int ebase, int it, int remaining,
                     const uint8_t lt_floor[LANES], int clamp_mins) {
    int d = row_deg(row);
    int odd = d & 1;
    int lut_it = imin(it, 6) - 1; /* switch (nb_iteration - nombre_iterations): 1..5, default = it6 */
    for (int f = 0; f < LANES; ++f) {
        int v[LDPC_MAXDEG], t[LDPC_MAXDEG], sg[LDPC_MAXDEG];
        int sign = 0, min1 = SAT_POS_VAR, min2 = SAT_POS_VAR;
        int eef = c->ef_elimination >= 1 && remaining <= c->ef_floor_iter_thresh && lt_floor[f] && w->chk[row][f]; /* :712-720 */
        for (int j = 0; j < d; ++j) {
            int n = row_vn(row, j);
            int l = w->L[n][f];
            int x = imin(imax(sat8(l - w->msg[ebase + j][f]), SAT_NEG_VAR), SAT_POS_VAR); /* :671-672 */
            /* EF_ELIMINATION 2 (:673-680): a regular VN whose checks are ALL unsatisfied sends an erasure the first time it
             * is visited in the iteration.  Only CDecoder_FAID.cpp has this block: in CDecoder_FAID_2B1C.cpp (clamp_mins)
             * mode 2 merely selects the (20, 6) thresholds at :120-123 */
            if (c->ef_elimination == 2 && !clamp_mins && vn_weight(n) == c->regular_col_weight && remaining <= c->ef_floor_iter_thresh &&
                w->vote[n][f] >= c->regular_col_weight && lt_floor[f] && !w->era[n][f]) {
                x = 0;
                w->era[n][f] = 1;
            }
            int sx = (x == 0) ? sat8(x + l) : x; /* :681 FAID2_SIGN_BACKTRACK */
            sg[j] = sx < 0;
            sign ^= sg[j];
            int a = iabs(x);
            v[j] = x;
            int wc = wclass(vn_weight(n));      /* :692-705 */
            int idx = a >= SAT_POS_MSG + 1 ? SAT_POS_MSG : a; /* :709-781 step map, :783-852 overflow case */
            int tm = eef ? c->v2c_lut_ef[lut_it][wc][idx] : c->v2c_lut[lut_it][wc][idx];
            t[j] = tm;
            int old = min1;
            min1 = imin(tm, min1);             /* :855 */
            min2 = imin(min2, imax(old, tm));  /* :856 */
        }
        if (clamp_mins) { /* CDecoder_FAID_2B1C.cpp:671-672 */
            min1 = imin(min1, SAT_POS_MSG);
            min2 = imin(min2, SAT_POS_MSG);
        }
        int c1 = imin(sat8(min2 - 0), SAT_POS_MSG); /* :864-865, offset = 0 (:9) */
        int c2 = imin(sat8(min1 - 0), SAT_POS_MSG);
        sign ^= odd;
        for (int j = 0; j < d; ++j) {
            int n = row_vn(row, j);
            int mag = (iabs(t[j]) == min1) ? c1 : c2; /* :910-914 */
            int neg = sign ^ sg[j];                    /* :915 */
            int m = neg ? -mag : mag;
            int l = imin(imax(sat8(v[j] + m), SAT_NEG_VAR), SAT_POS_VAR); /* :917-920: pre-LUT v */
            w->msg[ebase + j][f] = (int8_t)m;
            w->L[n][f] = (int8_t)l;
        }
    }
}

/* BF / DTBF / 2B1C post-processing.  CDecoder_OMSBF.cpp:2959-3511, CDecoder_FAID.cpp:6411-7088,
 * CDecoder_OMS_DTBF.cpp:2968-3650, CDecoder_FAID_2B1C.cpp:6124-6813.  Returns BFiter. */
static int bf_stage(work_t* w, const ldpc_b200_config* c) {
    int mode = c->bf_mode;
    for (int n = 0; n < N; ++n)
        for (int f = 0; f < LANES; ++f) {
            int l = w->L[n][f];
            w->hard[n][f] = l > 0;
            w->hch[n][f] = w->hard[n][f];
            w->hard2[n][f] = (l >= c->hard2_threshold) || (l <= -c->hard2_threshold);
            w->rec[n][f] = 0; /* mask_flip_record = {0} */
        }
    int t[LANES], Th[LANES], l0[LANES], l1[LANES];
    for (int f = 0; f < LANES; ++f) { t[f] = 1; Th[f] = c->regular_col_weight; l0[f] = 0; l1[f] = 0; }
    int L0 = (int8_t)c->dtbf_L0, L1 = (int8_t)c->dtbf_L1;
    int BFiter = 0;
    while (BFiter < c->bf_max_iter) {
        int err[LANES], maxv[LANES];
        memset(w->fv, 0, sizeof w->fv);
        for (int f = 0; f < LANES; ++f) { err[f] = 0; maxv[f] = 1; }
        for (int row = 0; row < M; ++row) {
            int d = row_deg(row);
            for (int f = 0; f < LANES; ++f) {
                int x = 0;
                for (int j = 0; j < d; ++j) x ^= w->hard[row_vn(row, j)][f];
                if (x) {
                    err[f] = usat8(err[f] + 1);
                    for (int j = 0; j < d; ++j) {
                        int n = row_vn(row, j);
                        w->fv[n][f] = (uint8_t)usat8(w->fv[n][f] + 1);
                    }
                }
            }
        }
        int any = 0;
        for (int f = 0; f < LANES; ++f) any |= err[f] > 0;
        if (!any) break; /* group-level break */
        if (mode == LDPC_B200_BF_PLAIN) {
            for (int n = 0; n < N; ++n)
                for (int f = 0; f < LANES; ++f) maxv[f] = imax(maxv[f], (int8_t)w->fv[n][f]); /* VECTOR_MAX is signed */
            for (int n = 0; n < N; ++n)
                for (int f = 0; f < LANES; ++f)
                    w->hard[n][f] ^= ((int8_t)w->fv[n][f] >= imin(maxv[f], 5)); /* CDecoder_OMSBF.cpp:3327-3335 */
        } else {
            for (int f = 0; f < LANES; ++f) {
                /* CDecoder_FAID.cpp:6787-6799 */
                if (!t[f]) Th[f] = sat8(Th[f] - c->dtbf_delta);
                int mx = t[f] && (l0[f] < L0);
                if (mx) { Th[f] = c->regular_col_weight + c->dtbf_alpha; l0[f] = sat8(l0[f] + 1); }
                int sub = t[f] && !mx && (l1[f] < L1);
                if (sub) { Th[f] = c->regular_col_weight + c->dtbf_alpha - c->dtbf_delta; l1[f] = sat8(l1[f] + 1); }
                int ssub = t[f] && !mx && !sub;
                if (ssub) Th[f] = c->regular_col_weight + c->dtbf_alpha - 2 * c->dtbf_delta;
                Th[f] = imax(Th[f], 1);
                t[f] = 0;
            }
            for (int n = 0; n < N; ++n) {
                if (vn_weight(n) != c->regular_col_weight) continue; /* :6808 */
                for (int f = 0; f < LANES; ++f) {
                    int differs = w->hard[n][f] ^ w->hch[n][f];
                    int vote = (int8_t)w->fv[n][f];
                    if (differs) vote = sat8(vote + c->dtbf_alpha); /* VECTOR_ADD_MASK: signed saturating */
                    int flip = vote >= Th[f];
                    w->rec[n][f] = (int8_t)flip;
                    t[f] |= flip;
                }
            }
            if (mode == LDPC_B200_BF_DTBF) {
                for (int n = 0; n < N; ++n)
                    for (int f = 0; f < LANES; ++f) w->hard[n][f] ^= w->rec[n][f]; /* :7084-7086 */
            } else { /* 2B1C, CDecoder_FAID_2B1C.cpp:6801-6814 */
                for (int n = 0; n < N; ++n)
                    for (int f = 0; f < LANES; ++f) {
                        int big = Th[f] >= c->regular_col_weight;
                        int r = w->rec[n][f];
                        if (big) {
                            w->hard[n][f] ^= r;
                            w->hard2[n][f] ^= r;
                        } else if (r) {
                            if (w->hard2[n][f]) w->hard2[n][f] = 0;
                            else w->hard[n][f] ^= 1;
                        }
                    }
            }
        }
        BFiter++;
    }
    return BFiter;
}

int ldpc_oracle_decode(const ldpc_b200_config* c, const int8_t* fixInput, int8_t* decodedBits, ldpc_oracle_info* info) {
    int method = c->decode_method;
    if (method < 0 || method > 5) method = 0; /* CSimulate.cpp:161-163 */
    work_t* w = (work_t*)malloc(sizeof(work_t));
    if (!w) return -1;
    memset(w->msg, 0, sizeof w->msg); /* CLDPC.cpp:230-232 */
    /* CLDPC.cpp:234-258: info region [f][j], then parity region [f][j]; lane f = frame f (CTool.cpp:9-289) */
    for (int f = 0; f < LANES; ++f) {
        for (int j = 0; j < K; ++j) w->L[j][f] = fixInput[f * K + j];
        for (int j = 0; j < M; ++j) w->L[K + j][f] = fixInput[LANES * K + f * M + j];
    }
    for (int i = 0; i < c->puncture_tail; ++i) /* CLDPC.cpp:270-272 */
        for (int f = 0; f < LANES; ++f) w->L[N - 1 - i][f] = 0;
    memset(w->chk, 0, sizeof w->chk);

    ldpc_oracle_info li;
    memset(&li, 0, sizeof li);
    for (int f = 0; f < LANES; ++f) li.conv_iter[f] = -1;

    int is_faid = (method == 2 || method == 5);
    int nombre_iterations = c->max_iteration;
    uint8_t lt_floor[LANES];
    memset(lt_floor, 0, sizeof lt_floor);
    while (nombre_iterations--) {
        int it = c->max_iteration - nombre_iterations; /* 1-based */
        if (method != 0) {
            int err[LANES];
            /* OMS family: unsigned saturation (<=255), FAID family: signed saturation (<=127) */
            syndrome_llr(w, err, is_faid ? 127 : 255);
            int any = 0;
            for (int f = 0; f < LANES; ++f) {
                any |= err[f] > 0;
                if (err[f] == 0 && li.conv_iter[f] < 0) li.conv_iter[f] = it - 1;
            }
            if (!any) break; /* group early stop: CDecoder_OMS.cpp:325-327, CDecoder_FAID.cpp:616-618 */
            for (int f = 0; f < LANES; ++f) {
                /* :328 VECTOR_LTU_MASK(error_sum, 100) ; FAID :619 VECTOR_LT_MASK(error_sum, floor_err_count) signed */
                lt_floor[f] = is_faid ? (err[f] < (int8_t)c->ef_floor_err_count)
                                      : ((unsigned)err[f] < (unsigned)(uint8_t)c->oms_floor_err_count);
                if (li.errsum_n < 64) li.errsum_log[li.errsum_n][f] = (uint8_t)err[f];
            }
            li.errsum_n++;
        }
        int ebase = 0;
        for (int row = 0; row < M; ++row) {
            switch (method) {
            case 0: row_nms(w, c, row, ebase); break;
            case 1: case 3: case 4: row_oms(w, c, row, ebase, nombre_iterations, lt_floor); break;
            default: row_faid(w, c, row, ebase, it, nombre_iterations, lt_floor, method == 5); break;
            }
            ebase += row_deg(row);
        }
        li.iters_executed++;
    }
    if (c->bf_mode != LDPC_B200_BF_NONE && method != 0 && method != 1) {
        li.bf_iters = bf_stage(w, c);
        for (int f = 0; f < LANES; ++f)
            for (int n = 0; n < N; ++n) decodedBits[f * N + n] = w->hard[n][f]; /* CDecoder_FAID.cpp:7091-7102 */
    } else {
        for (int f = 0; f < LANES; ++f)
            for (int n = 0; n < N; ++n) decodedBits[f * N + n] = w->L[n][f] > 0; /* CLDPC.cpp:2268-2270, CTool.cpp:291 */
    }
    if (info) *info = li;
    free(w);
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* Frame generation                                                                                  */
/* ------------------------------------------------------------------------------------------------ */

/* CLDPC.cpp:4524-4582.  _mm256_cvttps_epi32 returns INT_MIN for NaN / out-of-range ("integer indefinite"),
 * which the saturating packs turn into -128 and the clamp into -7. */
void ldpc_oracle_quantize_4bit(int8_t* out, const float* in, float scale, int64_t length) {
    for (int64_t i = 0; i < length; ++i) {
        float p = in[i] * scale;
        int q;
        if (!(p >= -2147483648.0f && p < 2147483648.0f)) q = -128; /* indefinite -> INT_MIN -> packs -> -128 */
        else {
            int32_t t = (int32_t)p; /* truncation toward zero */
            q = t > 127 ? 127 : (t < -128 ? -128 : t); /* packs_epi32 + packs_epi16 */
        }
        q = q > 7 ? 7 : q;
        q = q < -7 ? -7 : q;
        out[i] = (int8_t)q;
    }
}

/* The other quantisers of the reference (CLDPC.cpp:4385-4522, 4584-4770), vector path (length % 16 == 0 as in
 * CSimulate.cpp:124-132): 6 bit rounds to nearest even (_mm256_cvtps_epi32, :4436) and clamps to [-31,31]; 5 / 3 / 2 bit
 * truncate and clamp to [-16,15] / [-4,3] / [-2,1] (:4472-4473, :4592-4593, :4651-4652); 1 bit is +31 for q > 0 else -31
 * (:4746-4757).  Returns -1 for an unknown width. */
int ldpc_oracle_quantize_bits(int8_t* out, const float* in, float scale, int64_t length, int bits) {
    int lo, hi;
    switch (bits) {
    case 6: lo = -31; hi = 31; break;
    case 5: lo = -16; hi = 15; break;
    case 4: lo = -7; hi = 7; break;
    case 3: lo = -4; hi = 3; break;
    case 2: lo = -2; hi = 1; break;
    case 1: lo = -31; hi = 31; break;
    default: return -1;
    }
    for (int64_t i = 0; i < length; ++i) {
        float p = in[i] * scale;
        int q;
        if (!(p >= -2147483648.0f && p < 2147483648.0f)) q = -128;
        else {
            int32_t t = bits == 6 ? (int32_t)lrintf(p) /* default MXCSR rounding: nearest even */ : (int32_t)p;
            q = t > 127 ? 127 : (t < -128 ? -128 : t);
        }
        if (bits == 1) q = q > 0 ? 63 : -64;
        q = q > hi ? hi : q;
        q = q < lo ? lo : q;
        out[i] = (int8_t)q;
    }
    return 0;
}

void ldpc_oracle_transpose(const int8_t* src, int8_t* dst, int n) {
    for (int f = 0; f < 32; ++f)
        for (int i = 0; i < n; ++i) dst[i * 32 + f] = src[f * n + i];
}
void ldpc_oracle_itranspose(const int8_t* src, int8_t* dst, int n) {
    for (int f = 0; f < 32; ++f)
        for (int i = 0; i < n; ++i) dst[f * n + i] = src[i * 32 + f] > 0; /* CTool.cpp:291 */
}

/* CModulate.cpp:4-7 */
static const float table_256qam[16] = {-0.383482f, -0.536875f, -0.230089f, -0.076696f, -0.843661f, -0.690268f, -0.997054f, -1.150447f,
                                       0.383482f, 0.536875f, 0.230089f, 0.076696f, 0.843661f, 0.690268f, 0.997054f, 1.150447f};
static const float table_qpsk[2] = {-0.707107f, 0.707107f};
static const float table_16qam[4] = {-0.316228f, -0.948683f, 0.316228f, 0.948683f};
static const float table_64qam[8] = {-0.462910f, -0.154303f, -0.771517f, -1.08012f, 0.462910f, 0.154303f, 0.771517f, 1.08012f};

int ldpc_oracle_modulate(const int8_t* outputBits, int mod_type, int I, float* symbols) {
    const float* tab = mod_type == 2 ? table_qpsk : mod_type == 4 ? table_16qam : mod_type == 6 ? table_64qam : mod_type == 8 ? table_256qam : NULL;
    if (!tab || I < 1 || N % I) return -1;
    int8_t* il = (int8_t*)malloc((size_t)32 * N);
    int8_t* seq = (int8_t*)malloc((size_t)32 * N);
    /* CModulate.cpp:100-122: back to [frame][info|parity] */
    for (int f = 0; f < 32; ++f) {
        for (int j = 0; j < K; ++j) il[f * N + j] = outputBits[f * K + j];
        for (int j = 0; j < M; ++j) il[f * N + K + j] = outputBits[32 * K + f * M + j];
    }
    /* :137-149 per-frame interleaver */
    int64_t k = 0;
    for (int m = 0; m < 32; ++m)
        for (int j = 0; j < N / I; ++j)
            for (int i = 0; i < I; ++i) seq[k++] = il[N / I * i + j + N * m];
    /* :243-262 */
    int half = mod_type / 2;
    int64_t nsym = (int64_t)32 * N / mod_type;
    for (int64_t s = 0; s < nsym; ++s) {
        unsigned ti = 0, tq = 0;
        for (int j = 0; j < half; ++j) {
            ti += (unsigned)seq[s * mod_type + 2 * j] << (half - j - 1);
            tq += (unsigned)seq[s * mod_type + 2 * j + 1] << (half - j - 1);
        }
        symbols[2 * s] = tab[ti];
        symbols[2 * s + 1] = tab[tq];
    }
    free(il);
    free(seq);
    return 0;
}

int ldpc_oracle_demodulate(const float* symbols, int mod_type, int I, float* demod_out, float* deint) {
    if (!(mod_type == 2 || mod_type == 4 || mod_type == 6 || mod_type == 8) || I < 1 || N % I) return -1;
    float* demod = (float*)malloc(sizeof(float) * 32 * N);
    float* dil = (float*)malloc(sizeof(float) * 32 * N);
    int64_t nsym = (int64_t)32 * N / mod_type;
    /* CModulate.cpp:270-336: fabs(float) - double constant, evaluated in double, stored as float */
    for (int64_t s = 0; s < nsym; ++s) {
        float re = symbols[2 * s], im = symbols[2 * s + 1];
        float* d = demod + s * mod_type;
        d[0] = re;
        d[1] = im;
        if (mod_type == 4) {
            d[2] = (float)(fabs((double)re) - 0.6324555);
            d[3] = (float)(fabs((double)im) - 0.6324555);
        } else if (mod_type == 6) {
            d[2] = (float)(fabs((double)d[0]) - 0.6172134);
            d[3] = (float)(fabs((double)d[1]) - 0.6172134);
            d[4] = (float)(fabs((double)d[2]) - 0.3086067);
            d[5] = (float)(fabs((double)d[3]) - 0.3086067);
        } else if (mod_type == 8) { /* CModulate.cpp:340-356 */
            d[2] = (float)(fabs((double)d[0]) - 0.613568);
            d[3] = (float)(fabs((double)d[1]) - 0.613568);
            d[4] = (float)(fabs((double)d[2]) - 0.306784);
            d[5] = (float)(fabs((double)d[3]) - 0.306784);
            d[6] = (float)(fabs((double)d[4]) - 0.153392);
            d[7] = (float)(fabs((double)d[5]) - 0.153392);
        }
    }
    /* :161-172 */
    int64_t k = 0;
    for (int m = 0; m < 32; ++m)
        for (int j = 0; j < I; ++j)
            for (int i = 0; i < N / I; ++i) dil[k++] = demod[I * i + j + N * m];
    /* :176-202 */
    for (int f = 0; f < 32; ++f) {
        for (int j = 0; j < K; ++j) deint[f * K + j] = dil[N * f + j];
        for (int j = 0; j < M; ++j) deint[32 * K + f * M + j] = dil[f * N + K + j];
    }
    if (demod_out) memcpy(demod_out, demod, sizeof(float) * 32 * N);
    free(demod);
    free(dil);
    return 0;
}

/* CChannel.cpp:71-80 */
static float lcg_uniform(uint64_t st[3]) {
    float temp;
    st[0] = (st[0] * 249) % 61967;
    st[1] = (st[1] * 251) % 63443;
    st[2] = (st[2] * 252) % 63599;
    temp = (((float)st[0]) / ((float)61967)) + (((float)st[1]) / ((float)63443)) + (((float)st[2]) / ((float)63599));
    temp -= (int)temp;
    return temp;
}
/* CChannel.cpp:82-89 */
static float lcg_norm(double sigma, uint64_t st[3]) {
    float u1 = lcg_uniform(st);
    float u2 = lcg_uniform(st);
    float u = sigma * cos(2 * 3.1415926535897932384626433832795 * u2) * sqrt(-2.0 * log(1.0 - u1));
    return u;
}
/* CChannel.cpp:90-97; `sigma` here is what CSimulate passes: (float)(sigma / sqrt(2)) (CSimulate.cpp:126) */
void ldpc_oracle_awgn(const float* in, float* out, int64_t nsym, float sigma, uint64_t st[3]) {
    for (int64_t i = 0; i < nsym; ++i) {
        out[2 * i] = lcg_norm(sigma, st) + in[2 * i];
        out[2 * i + 1] = lcg_norm(sigma, st) + in[2 * i + 1];
    }
}
/* CSimulate.cpp:67-75 */
/* CModulate::BPSKModulation (CModulate.cpp:363-370): x = 2 b - 1 on the two-region outputBits buffer as it lies (no
 * interleaver, no regrouping); the received amplitude x + noise is quantised directly (CSimulate.cpp:121-124), so the
 * "LLR" keeps the two-region layout of fixInput. */
void ldpc_oracle_bpsk_modulate(const int8_t* outputBits, float* symbols) {
    for (int64_t i = 0; i < (int64_t)32 * N; ++i) symbols[i] = (float)(2 * outputBits[i] - 1);
}

float ldpc_oracle_sigma(float ebn0_db, int mod_type, double rate) {
    if (mod_type == 1) return (float)(1.0 / sqrt(2.0 * rate * mod_type * pow(10.0, 0.1 * ebn0_db)));
    return (float)(1.0 / sqrt(rate * mod_type * pow(10.0, 0.1 * ebn0_db)));
}

/* ------------------------------------------------------------------------------------------------ */
/* Encoder, syndrome, error counting                                                                 */
/* ------------------------------------------------------------------------------------------------ */

int ldpc_oracle_syndrome_weight(const int8_t* cw) {
    int wgt = 0;
    for (int row = 0; row < M; ++row) {
        int x = 0, d = row_deg(row);
        for (int j = 0; j < d; ++j) x ^= cw[row_vn(row, j)] & 1;
        wgt += x;
    }
    return wgt;
}

/* p = Hp^-1 (Hs s): the unique systematic codeword (parity part of H has full rank; tools/gen_code_tables.py). */
void ldpc_oracle_encode_frame(const int8_t* info, int8_t* cw) {
    uint8_t t[M];
    memcpy(cw, info, K);
    for (int row = 0; row < M; ++row) {
        int x = 0, d = row_deg(row);
        for (int j = 0; j < d; ++j) {
            int n = row_vn(row, j);
            if (n < K) x ^= info[n] & 1;
        }
        t[row] = (uint8_t)x;
    }
    for (int i = 0; i < LDPC_MB; ++i)
        for (int r = 0; r < QZ; ++r) {
            int x = 0;
            for (int j = 0; j < LDPC_MB; ++j)
                for (int cc = 0; cc < QZ; ++cc) {
                    int k = (r - cc) & (QZ - 1);
                    if ((ldpc_hpinv[i][j][k >> 5] >> (k & 31)) & 1) x ^= t[j * QZ + cc];
                }
            cw[K + i * QZ + r] = (int8_t)x;
        }
}

void ldpc_oracle_encode_group(const int8_t* inputBits, int8_t* outputBits) {
    int8_t cw[N];
    for (int f = 0; f < 32; ++f) {
        ldpc_oracle_encode_frame(inputBits + (size_t)f * K, cw);
        memcpy(outputBits + (size_t)f * K, cw, K);
        memcpy(outputBits + (size_t)32 * K + (size_t)f * M, cw + K, M);
    }
}

/* CLDPC.cpp:4842-4876: info bits only; parity mismatches are located but never counted */
void ldpc_oracle_calc_errors(const int8_t* inputBits, const int8_t* decodedBits, uint64_t stats3[3]) {
    stats3[0] = stats3[1] = stats3[2] = 0;
    for (int f = 0; f < 32; ++f) {
        uint64_t eb = 0;
        for (int j = 0; j < K; ++j) eb += decodedBits[(size_t)f * N + j] != inputBits[(size_t)f * K + j];
        if (eb > 0) {
            stats3[1] += eb;
            stats3[0]++;
            if (eb < 3) stats3[2]++;
        }
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* CPU baseline ("port")                                                                             */
/* ------------------------------------------------------------------------------------------------ */

typedef struct {
    const ldpc_b200_config* cfg;
    const int8_t* groups;
    int n_groups, tid;
    double min_seconds;
    int64_t frames;
} bench_arg;

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static void* bench_worker(void* p) {
    bench_arg* a = (bench_arg*)p;
    int8_t* out = (int8_t*)malloc((size_t)32 * N);
    double t0 = now_s();
    int g = a->tid % a->n_groups;
    a->frames = 0;
    do {
        ldpc_oracle_decode(a->cfg, a->groups + (size_t)g * 32 * N, out, NULL);
        a->frames += 32;
        g = (g + 1) % a->n_groups;
    } while (now_s() - t0 < a->min_seconds);
    free(out);
    return NULL;
}

double ldpc_oracle_bench_decode(const ldpc_b200_config* cfg, int n_threads, double min_seconds, const int8_t* groups,
                                int n_groups, int64_t* frames_done) {
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * n_threads);
    bench_arg* args = (bench_arg*)malloc(sizeof(bench_arg) * n_threads);
    double t0 = now_s();
    for (int t = 0; t < n_threads; ++t) {
        args[t] = (bench_arg){cfg, groups, n_groups, t, min_seconds, 0};
        pthread_create(&th[t], NULL, bench_worker, &args[t]);
    }
    int64_t total = 0;
    for (int t = 0; t < n_threads; ++t) {
        pthread_join(th[t], NULL);
        total += args[t].frames;
    }
    double el = now_s() - t0;
    free(th);
    free(args);
    if (frames_done) *frames_done = total;
    return total / el;
}
