/*
 * gcb_oracle.c -- CPU restatement of the reference algorithms (see gcb_oracle.h
 * for scope, usage rules and the parity-pinning statement).
 *
 * TEST INFRASTRUCTURE ONLY: checker and CPU baseline, never shipped.
 *
 * Written as plain sequential C that follows the Go loops one to one (same
 * gate order, same six/two hash calls per AND gate, same bit-by-bit
 * createLabels) so that timing it is a fair stand-in for the Go/AES-NI path:
 * with -maes the block cipher is the same AESENC chain Go's crypto/aes issues.
 */
#include "gcb_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#if defined(__AES__) && defined(__SSSE3__)
#include <immintrin.h>
#define ORC_AESNI 1
#else
#define ORC_AESNI 0
#endif

/* ======================================================================== */
/* AES, FIPS-197.  Portable byte-wise path + AES-NI path sharing one key     */
/* schedule.  The S-box is derived at start-up from the GF(2^8) inverse and   */
/* affine map rather than typed in.                                           */
/* ======================================================================== */

static uint8_t SBOX[256];
static int sbox_ready = 0;
static int use_aesni = ORC_AESNI;

static uint8_t gf_mul(uint8_t a, uint8_t b) {
    uint8_t p = 0;
    for (int i = 0; i < 8; i++) {
        if (b & 1) p ^= a;
        uint8_t hi = a & 0x80;
        a <<= 1;
        if (hi) a ^= 0x1b;
        b >>= 1;
    }
    return p;
}

static void sbox_init(void) {
    if (sbox_ready) return;
    for (int x = 0; x < 256; x++) {
        uint8_t inv = 0;
        if (x) {
            for (int y = 1; y < 256; y++)
                if (gf_mul((uint8_t)x, (uint8_t)y) == 1) { inv = (uint8_t)y; break; }
        }
        uint8_t s = inv, r = inv;
        for (int i = 0; i < 4; i++) { r = (uint8_t)((r << 1) | (r >> 7)); s ^= r; }
        SBOX[x] = s ^ 0x63;
    }
    __atomic_store_n(&sbox_ready, 1, __ATOMIC_RELEASE);
}

typedef struct {
    int nr;                      /* 10 / 12 / 14 */
    uint8_t rk[15][16];          /* round keys as byte blocks */
#if ORC_AESNI
    __m128i rkx[15];
#endif
} aes_ctx;

static int aes_init(aes_ctx *c, const uint8_t *key, uint32_t keylen) {
    if (keylen != 16 && keylen != 24 && keylen != 32) return ORC_E_KEYLEN;
    if (!__atomic_load_n(&sbox_ready, __ATOMIC_ACQUIRE)) sbox_init();
    int nk = (int)keylen / 4;
    c->nr = nk + 6;
    uint8_t w[60][4];
    memcpy(w, key, keylen);
    uint8_t rcon = 1;
    for (int i = nk; i < 4 * (c->nr + 1); i++) {
        uint8_t t[4];
        memcpy(t, w[i - 1], 4);
        if (i % nk == 0) {
            uint8_t t0 = t[0];
            t[0] = SBOX[t[1]] ^ rcon; t[1] = SBOX[t[2]]; t[2] = SBOX[t[3]]; t[3] = SBOX[t0];
            rcon = gf_mul(rcon, 2);
        } else if (nk > 6 && i % nk == 4) {
            for (int k = 0; k < 4; k++) t[k] = SBOX[t[k]];
        }
        for (int k = 0; k < 4; k++) w[i][k] = w[i - nk][k] ^ t[k];
    }
    memcpy(c->rk, w, 16 * (size_t)(c->nr + 1));
#if ORC_AESNI
    for (int r = 0; r <= c->nr; r++) c->rkx[r] = _mm_loadu_si128((const __m128i *)c->rk[r]);
#endif
    return ORC_OK;
}

static void aes_encrypt_portable(const aes_ctx *c, const uint8_t in[16], uint8_t out[16]) {
    uint8_t s[16], t[16];
    for (int i = 0; i < 16; i++) s[i] = in[i] ^ c->rk[0][i];
    for (int r = 1; r <= c->nr; r++) {
        /* SubBytes + ShiftRows: state is column-major, byte i = row i%4, col i/4 */
        for (int col = 0; col < 4; col++)
            for (int row = 0; row < 4; row++)
                t[4 * col + row] = SBOX[s[4 * ((col + row) & 3) + row]];
        if (r < c->nr) {
            for (int col = 0; col < 4; col++) {
                uint8_t a0 = t[4 * col], a1 = t[4 * col + 1], a2 = t[4 * col + 2], a3 = t[4 * col + 3];
                s[4 * col + 0] = gf_mul(a0, 2) ^ gf_mul(a1, 3) ^ a2 ^ a3;
                s[4 * col + 1] = a0 ^ gf_mul(a1, 2) ^ gf_mul(a2, 3) ^ a3;
                s[4 * col + 2] = a0 ^ a1 ^ gf_mul(a2, 2) ^ gf_mul(a3, 3);
                s[4 * col + 3] = gf_mul(a0, 3) ^ a1 ^ a2 ^ gf_mul(a3, 2);
            }
        } else {
            memcpy(s, t, 16);
        }
        for (int i = 0; i < 16; i++) s[i] ^= c->rk[r][i];
    }
    memcpy(out, s, 16);
}

static inline void aes_encrypt(const aes_ctx *c, const uint8_t in[16], uint8_t out[16]) {
#if ORC_AESNI
    if (use_aesni) {
        __m128i x = _mm_xor_si128(_mm_loadu_si128((const __m128i *)in), c->rkx[0]);
        for (int r = 1; r < c->nr; r++) x = _mm_aesenc_si128(x, c->rkx[r]);
        x = _mm_aesenclast_si128(x, c->rkx[c->nr]);
        _mm_storeu_si128((__m128i *)out, x);
        return;
    }
#endif
    aes_encrypt_portable(c, in, out);
}

int orc_have_aesni(void) { return ORC_AESNI; }
void orc_set_aesni(int on) { use_aesni = on && ORC_AESNI; }

int orc_aes_encrypt_block(const uint8_t *key, uint32_t keylen, const uint8_t in[16], uint8_t out[16]) {
    aes_ctx c;
    int rc = aes_init(&c, key, keylen);
    if (rc) return rc;
    aes_encrypt(&c, in, out);
    return ORC_OK;
}

/* ======================================================================== */
/* Label algebra, ot/label.go                                                */
/* ======================================================================== */

static inline void put_be64(uint8_t *p, uint64_t v) {
    for (int i = 0; i < 8; i++) p[i] = (uint8_t)(v >> (56 - 8 * i));
}
static inline uint64_t get_be64(const uint8_t *p) {
    uint64_t v = 0;
    for (int i = 0; i < 8; i++) v = (v << 8) | p[i];
    return v;
}
/* GetData, ot/label.go:105-108 */
void orc_label_to_bytes(orc_label l, uint8_t out[16]) { put_be64(out, l.d0); put_be64(out + 8, l.d1); }
/* SetData, ot/label.go:111-114 */
orc_label orc_label_from_bytes(const uint8_t in[16]) {
    orc_label l = { get_be64(in), get_be64(in + 8) };
    return l;
}
static inline orc_label lxor(orc_label a, orc_label b) { a.d0 ^= b.d0; a.d1 ^= b.d1; return a; }
static inline int lbl_s(orc_label l) { return (int)(l.d0 >> 63); }              /* S(), :65-67 */
static inline int lbl_eq(orc_label a, orc_label b) { return a.d0 == b.d0 && a.d1 == b.d1; }
/* Mul2, ot/label.go:79-83 */
orc_label orc_label_mul2(orc_label l) { l.d0 = (l.d0 << 1) | (l.d1 >> 63); l.d1 <<= 1; return l; }
/* Mul4, ot/label.go:86-90 */
orc_label orc_label_mul4(orc_label l) { l.d0 = (l.d0 << 2) | (l.d1 >> 62); l.d1 <<= 2; return l; }
/* Bit(i), ot/label.go:129-141: i<64 -> D0>>i, else D1>>(i-64) */
static inline unsigned lbl_bit(orc_label l, int i) {
    return (unsigned)(((i > 63) ? (l.d1 >> (i - 64)) : (l.d0 >> i)) & 1);
}

/* encryptHalf, circuit/garble.go:104-136: K = 2x ^ i ; pi(K) ^ K */
static inline orc_label enc_half(const aes_ctx *alg, orc_label x, uint32_t i) {
    orc_label k = orc_label_mul2(x);
    k.d1 ^= (uint64_t)i;
    uint8_t data[16];
    orc_label_to_bytes(k, data);
    aes_encrypt(alg, data, data);
    return lxor(orc_label_from_bytes(data), k);
}
/* makeK, circuit/garble.go:75-83 */
static inline orc_label make_k(orc_label a, orc_label b, uint32_t t) {
    orc_label k = lxor(orc_label_mul2(a), orc_label_mul4(b));
    k.d1 ^= (uint64_t)t;
    return k;
}
/* encrypt, circuit/garble.go:40-56 */
static inline orc_label enc2(const aes_ctx *alg, orc_label a, orc_label b, orc_label c, uint32_t t) {
    orc_label k = make_k(a, b, t);
    uint8_t data[16];
    orc_label_to_bytes(k, data);
    aes_encrypt(alg, data, data);
    return lxor(lxor(orc_label_from_bytes(data), k), c);
}
/* decrypt, circuit/garble.go:58-73 */
static inline orc_label dec2(const aes_ctx *alg, orc_label a, orc_label b, uint32_t t, orc_label c) {
    orc_label k = make_k(a, b, t);
    uint8_t data[16];
    orc_label_to_bytes(k, data);
    aes_encrypt(alg, data, data);
    return lxor(lxor(c, orc_label_from_bytes(data)), k);
}
/* idx / idxUnary, circuit/garble.go:20-38 */
static inline int idx2(orc_label a, orc_label b) { return (lbl_s(a) << 1) | lbl_s(b); }

int orc_encrypt_half(const uint8_t *key, uint32_t keylen, orc_label x, uint32_t tweak, orc_label *out) {
    aes_ctx c; int rc = aes_init(&c, key, keylen); if (rc) return rc;
    *out = enc_half(&c, x, tweak); return ORC_OK;
}
int orc_encrypt(const uint8_t *key, uint32_t keylen, orc_label a, orc_label b, orc_label c, uint32_t t, orc_label *out) {
    aes_ctx x; int rc = aes_init(&x, key, keylen); if (rc) return rc;
    *out = enc2(&x, a, b, c, t); return ORC_OK;
}
int orc_decrypt(const uint8_t *key, uint32_t keylen, orc_label a, orc_label b, uint32_t t, orc_label c, orc_label *out) {
    aes_ctx x; int rc = aes_init(&x, key, keylen); if (rc) return rc;
    *out = dec2(&x, a, b, t, c); return ORC_OK;
}

/* ======================================================================== */
/* One gate, garbler side: the arithmetic shared by Gate.garbleInto           */
/* (circuit/garble.go:311-482) and Streaming.garbleGate                       */
/* (circuit/stream_garble.go:195-389).  Returns the table window              */
/* [start, start+count) exactly as those functions do.                        */
/* ======================================================================== */
static int garble_gate_math(const aes_ctx *alg, uint8_t op, orc_wire a, orc_wire b, orc_label r,
                            uint32_t *idp, orc_label table[4], orc_wire *cout, int *start, int *count) {
    orc_wire c;
    memset(&c, 0, sizeof c);
    *start = 0; *count = 0;
    switch (op) {
    case ORC_XOR: {                                  /* garble.go:331-340 */
        orc_label l0 = lxor(a.l0, b.l0);
        c.l0 = l0; c.l1 = lxor(l0, r);
        break;
    }
    case ORC_XNOR: {                                 /* garble.go:342-351 */
        orc_label l0 = lxor(a.l0, b.l0);
        c.l0 = lxor(l0, r); c.l1 = l0;
        break;
    }
    case ORC_AND: {                                  /* garble.go:353-395 */
        int pa = lbl_s(a.l0), pb = lbl_s(b.l0);
        uint32_t j0 = *idp, j1 = *idp + 1;
        *idp += 2;
        /* first half gate: the reference hashes a.L0 twice; kept as is */
        orc_label tg = lxor(enc_half(alg, a.l0, j0), enc_half(alg, a.l1, j0));
        if (pb) tg = lxor(tg, r);
        orc_label wg0 = enc_half(alg, a.l0, j0);
        if (pa) wg0 = lxor(wg0, tg);
        /* second half gate */
        orc_label te = lxor(enc_half(alg, b.l0, j1), enc_half(alg, b.l1, j1));
        te = lxor(te, a.l0);
        orc_label we0 = enc_half(alg, b.l0, j1);
        if (pb) we0 = lxor(lxor(we0, te), a.l0);
        orc_label l0 = lxor(wg0, we0);
        c.l0 = l0; c.l1 = lxor(l0, r);
        table[0] = tg; table[1] = te;
        *count = 2;
        break;
    }
    case ORC_OR: {                                   /* garble.go:412-444; c is still zero */
        uint32_t id = (*idp)++;
        table[idx2(a.l0, b.l0)] = enc2(alg, a.l0, b.l0, c.l0, id);
        table[idx2(a.l0, b.l1)] = enc2(alg, a.l0, b.l1, c.l1, id);
        table[idx2(a.l1, b.l0)] = enc2(alg, a.l1, b.l0, c.l1, id);
        table[idx2(a.l1, b.l1)] = enc2(alg, a.l1, b.l1, c.l1, id);
        int l0i = idx2(a.l0, b.l0);
        c.l0 = table[0]; c.l1 = table[0];
        if (l0i == 0) c.l1 = lxor(c.l1, r); else c.l0 = lxor(c.l0, r);
        for (int i = 0; i < 4; i++) table[i] = lxor(table[i], (i == l0i) ? c.l0 : c.l1);
        *start = 1; *count = 3;
        break;
    }
    case ORC_INV: {                                  /* garble.go:446-474 */
        orc_label zero = {0, 0};
        uint32_t id = (*idp)++;
        table[lbl_s(a.l0)] = enc2(alg, a.l0, zero, c.l1, id);
        table[lbl_s(a.l1)] = enc2(alg, a.l1, zero, c.l0, id);
        int l0i = lbl_s(a.l0);
        c.l0 = table[0]; c.l1 = table[0];
        if (l0i == 0) c.l0 = lxor(c.l0, r); else c.l1 = lxor(c.l1, r);
        for (int i = 0; i < 2; i++) table[i] = lxor(table[i], (i == l0i) ? c.l1 : c.l0);
        *start = 1; *count = 1;
        break;
    }
    default:
        return ORC_E_BADOP;
    }
    *cout = c;
    return ORC_OK;
}

/* One gate, evaluator side (circuit/eval.go:44-110 == stream_evaluator.go:364-430). */
static int eval_gate_math(const aes_ctx *alg, uint8_t op, orc_label a, orc_label b,
                          const orc_label *row, int nrow, uint32_t *idp, orc_label *out) {
    orc_label c = {0, 0};
    switch (op) {
    case ORC_XOR: case ORC_XNOR:
        *out = lxor(a, b);
        return ORC_OK;
    case ORC_AND: {
        if (nrow != 2) return ORC_E_AND_ROWS;
        int sa = lbl_s(a), sb = lbl_s(b);
        uint32_t j0 = *idp, j1 = *idp + 1;
        *idp += 2;
        orc_label wg = enc_half(alg, a, j0);
        if (sa) wg = lxor(wg, row[0]);
        orc_label we = enc_half(alg, b, j1);
        if (sb) we = lxor(lxor(we, row[1]), a);
        *out = lxor(wg, we);
        return ORC_OK;
    }
    case ORC_OR: {
        int index = idx2(a, b);
        if (index > 0) {
            index--;
            if (index >= nrow) return ORC_E_ROW_INDEX;
            c = row[index];
        }
        *out = dec2(alg, a, b, *idp, c);
        (*idp)++;
        return ORC_OK;
    }
    case ORC_INV: {
        orc_label zero = {0, 0};
        int index = lbl_s(a);
        if (index > 0) {
            index--;
            if (index >= nrow) return ORC_E_ROW_INDEX;
            c = row[index];
        }
        *out = dec2(alg, a, zero, *idp, c);
        (*idp)++;
        return ORC_OK;
    }
    default:
        return ORC_E_BADOP;
    }
}

/* ======================================================================== */
/* Circuit.Garble, circuit/garble.go:248-308                                  */
/* ======================================================================== */
int orc_garble(const orc_gate *gates, uint32_t ngates, uint32_t nwires, uint32_t ninputs,
               const uint8_t *key, uint32_t keylen, const uint8_t *rand,
               orc_label *r_out, orc_wire *wires, orc_label *slab, uint32_t *row_off) {
    (void)nwires;
    /* R: first 16 bytes of the reader, then SetS(true) (:253-258) */
    orc_label r = orc_label_from_bytes(rand);
    r.d0 |= 0x8000000000000000ULL;
    aes_ctx alg;
    int rc = aes_init(&alg, key, keylen);
    if (rc) return rc;
    /* input wires (:271-278, makeLabels :145-157) */
    for (uint32_t i = 0; i < ninputs; i++) {
        orc_label l0 = orc_label_from_bytes(rand + 16 * (size_t)(1 + i));
        wires[i].l0 = l0;
        wires[i].l1 = lxor(l0, r);
    }
    uint32_t id = 0, off = 0;
    orc_label table[4];
    for (uint32_t i = 0; i < ngates; i++) {
        const orc_gate *g = &gates[i];
        orc_wire a, b, c;
        memset(&b, 0, sizeof b);
        if (g->op > ORC_INV) return ORC_E_BADOP;
        if (g->op != ORC_INV) b = wires[g->in1];
        a = wires[g->in0];
        int start, count;
        rc = garble_gate_math(&alg, g->op, a, b, r, &id, table, &c, &start, &count);
        if (rc) return rc;
        wires[g->out] = c;
        if (row_off) row_off[i] = off;
        for (int k = 0; k < count; k++) slab[off + k] = table[start + k];
        off += (uint32_t)count;
    }
    if (row_off) row_off[ngates] = off;
    if (r_out) *r_out = r;
    return ORC_OK;
}

/* Circuit.Eval, circuit/eval.go:17-115 */
int orc_eval(const orc_gate *gates, uint32_t ngates, uint32_t nwires,
             const uint8_t *key, uint32_t keylen, orc_label *wires,
             const orc_label *slab, const uint32_t *row_off) {
    (void)nwires;
    aes_ctx alg;
    int rc = aes_init(&alg, key, keylen);
    if (rc) return rc;
    uint32_t id = 0, off = 0;
    static const int rows_of[5] = {0, 0, 2, 3, 1};
    for (uint32_t i = 0; i < ngates; i++) {
        const orc_gate *g = &gates[i];
        if (g->op > ORC_INV) return ORC_E_BADOP;
        orc_label a = wires[g->in0], b = {0, 0}, out;
        if (g->op != ORC_INV) b = wires[g->in1];
        const orc_label *row;
        int nrow;
        if (row_off) { row = slab + row_off[i]; nrow = (int)(row_off[i + 1] - row_off[i]); }
        else { row = slab + off; nrow = rows_of[g->op]; off += (uint32_t)nrow; }
        rc = eval_gate_math(&alg, g->op, a, b, row, nrow, &id, &out);
        if (rc) return rc;
        wires[g->out] = out;
    }
    return ORC_OK;
}

/* ---- batch drivers (CPU baseline): one orc_garble / orc_eval per instance -- */
typedef struct {
    const orc_gate *gates; uint32_t ngates, nwires, ninputs, noutputs;
    const uint8_t *keys; uint32_t keylen, key_stride;
    uint32_t lo, hi, rows;
    const uint8_t *rand; orc_label *r_out; orc_label *tables; orc_wire *io_wires;
    const orc_label *ctables; const orc_label *in_labels; orc_label *out_labels;
    int rc;
} batch_job;

static void *garble_worker(void *p) {
    batch_job *j = (batch_job *)p;
    orc_wire *wires = (orc_wire *)malloc(sizeof(orc_wire) * j->nwires);
    size_t rs = 16 * (size_t)(1 + j->ninputs);
    for (uint32_t i = j->lo; i < j->hi && !j->rc; i++) {
        const uint8_t *key = j->keys + (size_t)i * j->key_stride;
        j->rc = orc_garble(j->gates, j->ngates, j->nwires, j->ninputs, key, j->keylen,
                           j->rand + rs * i, j->r_out ? &j->r_out[i] : NULL, wires,
                           j->tables + (size_t)i * j->rows, NULL);
        if (j->io_wires) {
            orc_wire *io = j->io_wires + (size_t)i * (j->ninputs + j->noutputs);
            memcpy(io, wires, sizeof(orc_wire) * j->ninputs);
            memcpy(io + j->ninputs, wires + j->nwires - j->noutputs, sizeof(orc_wire) * j->noutputs);
        }
    }
    free(wires);
    return NULL;
}

static void *eval_worker(void *p) {
    batch_job *j = (batch_job *)p;
    orc_label *wires = (orc_label *)malloc(sizeof(orc_label) * j->nwires);
    for (uint32_t i = j->lo; i < j->hi && !j->rc; i++) {
        const uint8_t *key = j->keys + (size_t)i * j->key_stride;
        memcpy(wires, j->in_labels + (size_t)i * j->ninputs, sizeof(orc_label) * j->ninputs);
        j->rc = orc_eval(j->gates, j->ngates, j->nwires, key, j->keylen, wires,
                         j->ctables + (size_t)i * j->rows, NULL);
        memcpy(j->out_labels + (size_t)i * j->noutputs, wires + j->nwires - j->noutputs,
               sizeof(orc_label) * j->noutputs);
    }
    free(wires);
    return NULL;
}

static uint32_t count_rows(const orc_gate *gates, uint32_t ngates) {
    static const uint32_t rows_of[5] = {0, 0, 2, 3, 1};
    uint32_t n = 0;
    for (uint32_t i = 0; i < ngates; i++) n += gates[i].op <= ORC_INV ? rows_of[gates[i].op] : 0;
    return n;
}

static int run_batch(batch_job *proto, uint32_t batch, int threads, void *(*fn)(void *)) {
    if (threads < 1) threads = 1;
    if ((uint32_t)threads > batch) threads = (int)(batch ? batch : 1);
    pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    batch_job *jobs = (batch_job *)malloc(sizeof(batch_job) * (size_t)threads);
    for (int t = 0; t < threads; t++) {
        jobs[t] = *proto;
        jobs[t].lo = (uint32_t)((uint64_t)batch * (uint64_t)t / (uint64_t)threads);
        jobs[t].hi = (uint32_t)((uint64_t)batch * (uint64_t)(t + 1) / (uint64_t)threads);
        jobs[t].rc = 0;
        if (threads == 1) fn(&jobs[t]);
        else pthread_create(&tid[t], NULL, fn, &jobs[t]);
    }
    int rc = 0;
    for (int t = 0; t < threads; t++) {
        if (threads > 1) pthread_join(tid[t], NULL);
        if (jobs[t].rc && !rc) rc = jobs[t].rc;
    }
    free(tid); free(jobs);
    return rc;
}

int orc_garble_batch(const orc_gate *gates, uint32_t ngates, uint32_t nwires, uint32_t ninputs,
                     uint32_t noutputs, const uint8_t *keys, uint32_t keylen, uint32_t key_stride,
                     uint32_t batch, const uint8_t *rand, orc_label *r_out, orc_label *tables,
                     orc_wire *io_wires, int threads) {
    sbox_init();
    batch_job j;
    memset(&j, 0, sizeof j);
    j.gates = gates; j.ngates = ngates; j.nwires = nwires; j.ninputs = ninputs; j.noutputs = noutputs;
    j.keys = keys; j.keylen = keylen; j.key_stride = key_stride; j.rows = count_rows(gates, ngates);
    j.rand = rand; j.r_out = r_out; j.tables = tables; j.io_wires = io_wires;
    return run_batch(&j, batch, threads, garble_worker);
}

int orc_eval_batch(const orc_gate *gates, uint32_t ngates, uint32_t nwires, uint32_t ninputs,
                   uint32_t noutputs, const uint8_t *keys, uint32_t keylen, uint32_t key_stride,
                   uint32_t batch, const orc_label *tables, const orc_label *in_labels,
                   orc_label *out_labels, int threads) {
    sbox_init();
    batch_job j;
    memset(&j, 0, sizeof j);
    j.gates = gates; j.ngates = ngates; j.nwires = nwires; j.ninputs = ninputs; j.noutputs = noutputs;
    j.keys = keys; j.keylen = keylen; j.key_stride = key_stride; j.rows = count_rows(gates, ngates);
    j.ctables = tables; j.in_labels = in_labels; j.out_labels = out_labels;
    return run_batch(&j, batch, threads, eval_worker);
}

/* ======================================================================== */
/* Streaming garbler, circuit/stream_garble.go                                */
/* The reference pages the permanent wires in 64Ki blocks (:78-100); a flat    */
/* growable array is observationally identical.                               */
/* ======================================================================== */
struct orc_stream {
    aes_ctx alg;
    orc_label r;
    orc_wire *wires; size_t nwires;        /* permanent wires */
    orc_wire *tmp; size_t ntmp;
};

static void stream_ensure(orc_stream *s, uint32_t max) {          /* ensureWires :95-100 */
    size_t need = ((size_t)max / 0x10000 + 1) * 0x10000;
    if (need > s->nwires) {
        s->wires = (orc_wire *)realloc(s->wires, need * sizeof(orc_wire));
        memset(s->wires + s->nwires, 0, (need - s->nwires) * sizeof(orc_wire));
        s->nwires = need;
    }
}

orc_stream *orc_stream_new(const uint8_t *key, uint32_t keylen, const uint8_t *rand,
                           const uint32_t *input_ids, uint32_t ninputs) {
    orc_stream *s = (orc_stream *)calloc(1, sizeof *s);
    if (aes_init(&s->alg, key, keylen)) { free(s); return NULL; }
    s->r = orc_label_from_bytes(rand);                            /* :46-50 */
    s->r.d0 |= 0x8000000000000000ULL;
    uint32_t mx = 0;
    for (uint32_t i = 0; i < ninputs; i++) if (input_ids[i] > mx) mx = input_ids[i];
    stream_ensure(s, mx);
    for (uint32_t i = 0; i < ninputs; i++) {                      /* :66-73 */
        orc_label l0 = orc_label_from_bytes(rand + 16 * (size_t)(1 + i));
        s->wires[input_ids[i]].l0 = l0;
        s->wires[input_ids[i]].l1 = lxor(l0, s->r);
    }
    return s;
}
void orc_stream_free(orc_stream *s) { if (s) { free(s->wires); free(s->tmp); free(s); } }
orc_label orc_stream_r(const orc_stream *s) { return s->r; }
orc_wire orc_stream_get_input(const orc_stream *s, uint32_t id) { return s->wires[id]; }
void orc_stream_set_wire(orc_stream *s, uint32_t id, orc_wire w) { stream_ensure(s, id); s->wires[id] = w; }

static inline void put_be16(uint8_t *p, uint32_t v) { p[0] = (uint8_t)(v >> 8); p[1] = (uint8_t)v; }
static inline void put_be32(uint8_t *p, uint32_t v) {
    p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v;
}

int orc_stream_garble(orc_stream *s, const orc_gate *gates, uint32_t ngates, uint32_t nwires,
                      const uint32_t *in, uint32_t nin, const uint32_t *out, uint32_t nout,
                      uint8_t *buf, size_t cap, size_t *written) {
    /* initCircuit :102-114 */
    uint32_t mx = 0;
    for (uint32_t i = 0; i < nin; i++) if (in[i] > mx) mx = in[i];
    for (uint32_t i = 0; i < nout; i++) if (out[i] > mx) mx = out[i];
    stream_ensure(s, mx);
    if (s->ntmp < nwires) {
        free(s->tmp);
        s->tmp = (orc_wire *)calloc(nwires ? nwires : 1, sizeof(orc_wire));
        s->ntmp = nwires;
    }
    uint32_t first_tmp = nin, first_out = nwires - nout;
    uint32_t id = 0;                                              /* per call, :174 */
    size_t pos = 0;
    orc_label table[4];
    for (uint32_t gi = 0; gi < ngates; gi++) {
        const orc_gate *g = &gates[gi];
        if (g->op > ORC_INV) return ORC_E_BADOP;
        if (cap - pos < 64) return ORC_E_BUFFER;                  /* 1+12+48 = 61 bytes max */
        orc_wire a, b, c;
        uint32_t ai = 0, bi = 0, ci;
        int at = 0, bt = 0, ct = 0;
        memset(&b, 0, sizeof b);
        /* Get :131-141 */
#define STREAM_GET(W, WI, WT, X) do { uint32_t w_ = (X); \
        if (w_ < first_tmp) { WI = in[w_]; W = s->wires[WI]; WT = 0; } \
        else if (w_ >= first_out) { WI = out[w_ - first_out]; W = s->wires[WI]; WT = 0; } \
        else { WI = w_; W = s->tmp[w_]; WT = 1; } } while (0)
        if (g->op != ORC_INV) STREAM_GET(b, bi, bt, g->in1);
        STREAM_GET(a, ai, at, g->in0);
#undef STREAM_GET
        int start, count;
        int rc = garble_gate_math(&s->alg, g->op, a, b, s->r, &id, table, &c, &start, &count);
        if (rc) return rc;
        /* output placement :379-389 */
        if (g->out < first_tmp) { ci = in[g->out]; s->wires[ci] = c; }
        else if (g->out >= first_out) { ci = out[g->out - first_out]; s->wires[ci] = c; }
        else { ci = g->out; ct = 1; s->tmp[g->out] = c; }
        /* record header :391-440 */
        uint8_t op = g->op;
        if (at) op |= 0x80;
        if (bt) op |= 0x40;
        if (ct) op |= 0x20;
        int wire_count = (g->op == ORC_INV) ? 2 : 3;
        if (ai <= 0xffff && bi <= 0xffff && ci <= 0xffff) {
            buf[pos++] = op | 0x10;
            if (wire_count == 3) { put_be16(buf + pos, ai); put_be16(buf + pos + 2, bi); put_be16(buf + pos + 4, ci); pos += 6; }
            else { put_be16(buf + pos, ai); put_be16(buf + pos + 2, ci); pos += 4; }
        } else {
            buf[pos++] = op;
            if (wire_count == 3) { put_be32(buf + pos, ai); put_be32(buf + pos + 4, bi); put_be32(buf + pos + 8, ci); pos += 12; }
            else { put_be32(buf + pos, ai); put_be32(buf + pos + 4, ci); pos += 8; }
        }
        /* rows :442-446 */
        for (int k = 0; k < count; k++) { orc_label_to_bytes(table[start + k], buf + pos); pos += 16; }
    }
    *written = pos;
    return ORC_OK;
}

/* ======================================================================== */
/* Streaming evaluator, circuit/stream_evaluator.go:29-96, 270-432            */
/* ======================================================================== */
struct orc_seval {
    aes_ctx alg;
    orc_label *wires; size_t nwires;
    orc_label *tmp; size_t ntmp;
};
static void seval_ensure(orc_seval *s, size_t max) {
    size_t need = (max / 0x10000 + 1) * 0x10000;
    if (need > s->nwires) {
        s->wires = (orc_label *)realloc(s->wires, need * sizeof(orc_label));
        memset(s->wires + s->nwires, 0, (need - s->nwires) * sizeof(orc_label));
        s->nwires = need;
    }
}
orc_seval *orc_seval_new(const uint8_t *key, uint32_t keylen) {
    orc_seval *s = (orc_seval *)calloc(1, sizeof *s);
    if (aes_init(&s->alg, key, keylen)) { free(s); return NULL; }
    return s;
}
void orc_seval_free(orc_seval *s) { if (s) { free(s->wires); free(s->tmp); free(s); } }
void orc_seval_set(orc_seval *s, uint32_t id, orc_label l) { seval_ensure(s, id); s->wires[id] = l; }
orc_label orc_seval_get(const orc_seval *s, uint32_t id) { return s->wires[id]; }

int orc_seval_circuit(orc_seval *s, const uint8_t *buf, size_t len, uint32_t ngates,
                      uint32_t ntmp, uint32_t nwires, size_t *consumed) {
    seval_ensure(s, nwires);                                      /* InitCircuit :86-91 */
    if (s->ntmp < ntmp) {
        free(s->tmp);
        s->tmp = (orc_label *)calloc(ntmp ? ntmp : 1, sizeof(orc_label));
        s->ntmp = ntmp;
    }
    uint32_t id = 0;                                              /* :270 */
    size_t pos = 0;
    static const int rows_of[5] = {0, 0, 2, 3, 1};
    for (uint32_t i = 0; i < ngates; i++) {
        if (pos >= len) return ORC_E_BUFFER;
        uint8_t gop = buf[pos++];
        int at = (gop & 0x80) != 0, bt = (gop & 0x40) != 0, ct = (gop & 0x20) != 0;
        int sh = (gop & 0x10) != 0;
        gop &= 0x0f;
        if (gop > ORC_INV) return ORC_E_BADOP;
        int nidx = (gop == ORC_INV) ? 2 : 3, isz = sh ? 2 : 4;
        if (len - pos < (size_t)(nidx * isz + 16 * rows_of[gop])) return ORC_E_BUFFER;
        uint32_t ix[3] = {0, 0, 0};
        for (int k = 0; k < nidx; k++) {
            uint32_t v = 0;
            for (int q = 0; q < isz; q++) v = (v << 8) | buf[pos++];
            ix[k] = v;
        }
        uint32_t ai = ix[0], bi = (nidx == 3) ? ix[1] : 0, ci = ix[nidx - 1];
        orc_label rows[3];
        int nrow = rows_of[gop];
        for (int k = 0; k < nrow; k++) { rows[k] = orc_label_from_bytes(buf + pos); pos += 16; }
        orc_label a = at ? s->tmp[ai] : s->wires[ai], b = {0, 0}, o;
        if (gop != ORC_INV) b = bt ? s->tmp[bi] : s->wires[bi];
        int rc = eval_gate_math(&s->alg, gop, a, b, rows, nrow, &id, &o);
        if (rc) return rc;
        if (ct) s->tmp[ci] = o; else s->wires[ci] = o;
    }
    if (consumed) *consumed = pos;
    return ORC_OK;
}

/* ======================================================================== */
/* IKNP, ot/iknp.go                                                           */
/* ======================================================================== */
enum { IKNP_K = 128, CHUNK_SIZE = 8 * 1024, CHUNK_BYTE_ROWS = CHUNK_SIZE / IKNP_K,
       CHUNK_ROWS = CHUNK_BYTE_ROWS * 8 };                        /* :63-77 */

/* newPrg + prg, ot/iknp.go:622-637.  Go's cipher.NewCTR with a zero IV:
 * keystream block j = AES_k(j as a 128-bit big-endian integer).  `pos` is the
 * number of keystream bytes already consumed (the stream is stateful). */
static void prg_ctx(const aes_ctx *c, uint64_t pos, uint8_t *buf, size_t n) {
    uint8_t ctr[16], ks[16];
    while (n) {
        uint64_t blk = pos / 16;
        unsigned o = (unsigned)(pos % 16);
        memset(ctr, 0, 8);
        put_be64(ctr + 8, blk);
        aes_encrypt(c, ctr, ks);
        size_t take = 16 - o;
        if (take > n) take = n;
        memcpy(buf, ks + o, take);
        buf += take; n -= take; pos += take;
    }
}
void orc_prg(orc_label key, uint64_t pos, uint8_t *buf, size_t n) {
    uint8_t kb[16];
    aes_ctx c;
    orc_label_to_bytes(key, kb);
    aes_init(&c, kb, 16);
    prg_ctx(&c, pos, buf, n);
}

/* createLabels, ot/iknp.go:647-683 (same loop nest, bit by bit) */
void orc_create_labels(orc_label *l, size_t nl, const uint8_t *buf, int w) {
    size_t end = (size_t)w * 8;
    if (end > nl) end = nl;
    for (int row = 0; row < w; row++) {
        orc_label out[8];
        memset(out, 0, sizeof out);
        for (int j = 0; j < 128; j++) {
            uint8_t b = buf[j * w + row];
            uint64_t mask = (uint64_t)1 << ((unsigned)j & 63);
            for (int bit = 0; bit < 8; bit++) {
                if ((b >> bit) & 1) {
                    if (j < 64) out[bit].d0 |= mask; else out[bit].d1 |= mask;
                }
            }
        }
        size_t base = (size_t)row * 8;
        for (int bit = 0; bit < 8; bit++) {
            size_t i = base + (size_t)bit;
            if (i >= end) return;
            l[i] = out[bit];
        }
    }
}

static void prg_init_all(aes_ctx *ctx, const orc_label *keys, int n) {
    for (int i = 0; i < n; i++) {
        uint8_t kb[16];
        orc_label_to_bytes(keys[i], kb);
        aes_init(&ctx[i], kb, 16);
    }
}

/* IKNPReceiver.receive, ot/iknp.go:468-511 */
int orc_iknp_receive(const orc_label k0[128], const orc_label k1[128], uint64_t *pos,
                     const uint8_t *choice, uint64_t n, uint8_t *u_out, size_t u_cap,
                     size_t *u_len, orc_label *result) {
    aes_ctx g0[IKNP_K], g1[IKNP_K];
    prg_init_all(g0, k0, IKNP_K);
    prg_init_all(g1, k1, IKNP_K);
    size_t nb = (size_t)((n + 7) / 8);
    uint8_t *bbuf = (uint8_t *)calloc(nb + CHUNK_BYTE_ROWS, 1);
    for (uint64_t i = 0; i < n; i++) if (choice[i]) bbuf[i / 8] |= (uint8_t)(1u << (i % 8));
    uint8_t chunk[CHUNK_SIZE], out[CHUNK_SIZE], tmp[CHUNK_BYTE_ROWS];
    size_t ulen = 0;
    for (uint64_t ofs = 0; ofs < n;) {
        uint64_t rows = CHUNK_ROWS, avail = n - ofs;
        if (rows > avail) rows = avail;
        int byte_rows = (int)((rows + 7) / 8);
        for (int i = 0; i < IKNP_K; i++) {
            prg_ctx(&g0[i], *pos, chunk + i * byte_rows, (size_t)byte_rows);
            prg_ctx(&g1[i], *pos, tmp, (size_t)byte_rows);
            for (int r = 0; r < byte_rows; r++)
                out[i * byte_rows + r] = tmp[r] ^ chunk[i * byte_rows + r] ^ bbuf[ofs / 8 + (uint64_t)r];
        }
        *pos += (uint64_t)byte_rows;
        if (ulen + (size_t)byte_rows * 128 > u_cap) { free(bbuf); return ORC_E_BUFFER; }
        memcpy(u_out + ulen, out, (size_t)byte_rows * 128);
        ulen += (size_t)byte_rows * 128;
        orc_create_labels(result + ofs, (size_t)(n - ofs), chunk, byte_rows);
        ofs += rows;
    }
    free(bbuf);
    *u_len = ulen;
    return ORC_OK;
}

/* One sender chunk: ot/iknp.go:213-220 (shared by send and SendBits) */
static void sender_chunk(aes_ctx *g0, orc_label delta, uint64_t *pos, const uint8_t *chunk,
                         int byte_rows, uint8_t *t) {
    for (int i = 0; i < IKNP_K; i++) {
        prg_ctx(&g0[i], *pos, t + i * byte_rows, (size_t)byte_rows);
        if (lbl_bit(delta, i) == 1)
            for (int r = 0; r < byte_rows; r++) t[i * byte_rows + r] ^= chunk[i * byte_rows + r];
    }
    *pos += (uint64_t)byte_rows;
}

/* The receiver frames one SendData per chunk; the sender learns byteRows from
 * the frame length.  Here the frames are concatenated, so the chunk lengths are
 * re-derived the way the receiver produced them (:482-488). */
static int next_chunk_rows(uint64_t n, uint64_t ofs) {
    uint64_t rows = CHUNK_ROWS, avail = n - ofs;
    if (rows > avail) rows = avail;
    return (int)((rows + 7) / 8);
}

/* IKNPSender.send, ot/iknp.go:197-226 */
int orc_iknp_send(const orc_label k[128], orc_label delta, uint64_t *pos, const uint8_t *u,
                  size_t u_len, uint64_t n, orc_label *result) {
    aes_ctx g0[IKNP_K];
    prg_init_all(g0, k, IKNP_K);
    size_t upos = 0;
    for (uint64_t ofs = 0; ofs < n;) {
        int byte_rows = next_chunk_rows(n, ofs);
        if (upos + (size_t)byte_rows * 128 > u_len) return ORC_E_CHUNK;
        uint8_t t[CHUNK_SIZE];
        memset(t, 0, sizeof t);
        sender_chunk(g0, delta, pos, u + upos, byte_rows, t);
        upos += (size_t)byte_rows * 128;
        orc_create_labels(result + ofs, (size_t)(n - ofs), t, byte_rows);
        ofs += (uint64_t)byte_rows * 8;
    }
    return ORC_OK;
}

/* IKNPReceiver.ReceiveBits, ot/iknp.go:554-620.  Note (reference behaviour,
 * kept): the choice XOR runs over `words = byteRows/8` whole 64-bit words only
 * (:583,:592-596), so a final chunk whose byteRows is not a multiple of 8
 * leaves its trailing rows' choice bits out of U. */
int orc_iknp_receive_bits(const orc_label k0[128], const orc_label k1[128], uint64_t *pos,
                          const uint64_t *choices, uint64_t n, uint8_t *u_out, size_t u_cap,
                          size_t *u_len, uint64_t *result) {
    aes_ctx g0[IKNP_K], g1[IKNP_K];
    prg_init_all(g0, k0, IKNP_K);
    prg_init_all(g1, k1, IKNP_K);
    uint8_t chunk[CHUNK_SIZE], tmp[CHUNK_SIZE], ucol[CHUNK_SIZE];
    size_t ulen = 0;
    for (uint64_t ofs = 0; ofs < n;) {
        uint64_t rows = CHUNK_ROWS, avail = n - ofs;
        if (rows > avail) rows = avail;
        int byte_rows = (int)((rows + 7) / 8);
        uint64_t word_offset = ofs / 64;
        int words = byte_rows / 8;
        for (int i = 0; i < IKNP_K; i++) {
            prg_ctx(&g0[i], *pos, chunk + i * byte_rows, (size_t)byte_rows);
            prg_ctx(&g1[i], *pos, tmp, (size_t)byte_rows);
            for (int r = 0; r < byte_rows; r++) tmp[r] ^= chunk[i * byte_rows + r];
            for (int w = 0; w < words; w++) {
                uint64_t cw = choices[word_offset + (uint64_t)w];
                for (int q = 0; q < 8; q++) tmp[w * 8 + q] ^= (uint8_t)(cw >> (8 * q));
            }
            memcpy(ucol + i * byte_rows, tmp, (size_t)byte_rows);
        }
        *pos += (uint64_t)byte_rows;
        if (ulen + (size_t)byte_rows * 128 > u_cap) return ORC_E_BUFFER;
        memcpy(u_out + ulen, ucol, (size_t)byte_rows * 128);
        ulen += (size_t)byte_rows * 128;
        orc_label labels[CHUNK_ROWS];
        memset(labels, 0, sizeof labels);
        orc_create_labels(labels, CHUNK_ROWS, chunk, byte_rows);
        for (uint64_t row = 0; row < rows; row++)
            if (lbl_bit(labels[row], 0) == 1) {
                uint64_t ix = ofs + row;
                result[ix / 64] |= (uint64_t)1 << (ix % 64);
            }
        ofs += rows;
    }
    *u_len = ulen;
    return ORC_OK;
}

/* IKNPSender.SendBits, ot/iknp.go:259-310 */
int orc_iknp_send_bits(const orc_label k[128], orc_label delta, uint64_t *pos, const uint8_t *u,
                       size_t u_len, uint64_t n, uint64_t *result) {
    aes_ctx g0[IKNP_K];
    prg_init_all(g0, k, IKNP_K);
    size_t upos = 0;
    for (uint64_t ofs = 0; ofs < n;) {
        int byte_rows = next_chunk_rows(n, ofs);
        if (upos + (size_t)byte_rows * 128 > u_len) return ORC_E_CHUNK;
        uint8_t t[CHUNK_SIZE];
        memset(t, 0, sizeof t);
        sender_chunk(g0, delta, pos, u + upos, byte_rows, t);
        upos += (size_t)byte_rows * 128;
        uint64_t max_rows = (uint64_t)byte_rows * 8;
        if (max_rows > n - ofs) max_rows = n - ofs;
        for (uint64_t row = 0; row < max_rows; row++)
            if ((t[row / 8] >> (row % 8)) & 1) {
                uint64_t ix = ofs + row;
                result[ix / 64] |= (uint64_t)1 << (ix % 64);
            }
        ofs += max_rows;
    }
    return ORC_OK;
}

/* ======================================================================== */
/* MiTCCRH, ot/mitccrh.go                                                     */
/* ======================================================================== */
void orc_mitccrh_init(orc_mitccrh *m, orc_label seed, int batch_size) {          /* :61-68 */
    memset(m, 0, sizeof *m);
    m->batch_size = batch_size;
    m->start = seed;
    m->key_used = batch_size;              /* force renew on first use */
}
static void mitccrh_renew(orc_mitccrh *m) {                                       /* :70-89 */
    for (int i = 0; i < m->batch_size; i++) {
        orc_label key = { m->gid, 0 };
        m->gid++;
        key = lxor(key, m->start);
        orc_label_to_bytes(key, m->keys[i]);
    }
    m->key_used = 0;
}
int orc_mitccrh_hash(orc_mitccrh *m, orc_label *blks, int k, int h) {            /* :93-128 */
    if (k > m->batch_size || k <= 0 || m->batch_size % k != 0 || m->batch_size > 64) return ORC_E_ARG;
    if (m->key_used == m->batch_size) mitccrh_renew(m);
    for (int i = 0; i < k; i++) {
        aes_ctx c;
        aes_init(&c, m->keys[m->key_used + i], 16);
        for (int j = 0; j < h; j++) {
            uint8_t tmp[16];
            orc_label_to_bytes(blks[i * h + j], tmp);
            aes_encrypt(&c, tmp, tmp);
            blks[i * h + j] = lxor(blks[i * h + j], orc_label_from_bytes(tmp));
        }
    }
    m->key_used += k;
    return ORC_OK;
}

/* ======================================================================== */
/* COT / ROT post-processing, ot/cot.go:136-235 and ot/rot.go:132-202          */
/* (otBatchSize = 8, ot/cot.go:47).  The pad array persists across batches,    */
/* so the stale tail entries of a final partial batch are hashed too -- that   */
/* costs time but never reaches an output.                                    */
/* ======================================================================== */
enum { OT_BATCH = 8 };

void orc_cot_send(const orc_label *data, orc_label delta, orc_label seed, const orc_wire *wires,
                  uint64_t n, orc_label *out_msgs) {
    orc_mitccrh m;
    orc_mitccrh_init(&m, seed, OT_BATCH);
    orc_label pad[2 * OT_BATCH];
    memset(pad, 0, sizeof pad);
    for (uint64_t i = 0; i < n; i += OT_BATCH) {
        uint64_t end = i + OT_BATCH;
        if (end > n) end = n;
        for (uint64_t j = i; j < end; j++) {
            pad[2 * (j - i)] = data[j];
            pad[2 * (j - i) + 1] = lxor(data[j], delta);
        }
        orc_mitccrh_hash(&m, pad, OT_BATCH, 2);
        for (uint64_t j = i; j < end; j++) {
            pad[2 * (j - i)] = lxor(pad[2 * (j - i)], wires[j].l0);
            pad[2 * (j - i) + 1] = lxor(pad[2 * (j - i) + 1], wires[j].l1);
        }
        for (uint64_t j = 0; j < 2 * (end - i); j++) out_msgs[2 * i + j] = pad[j];
    }
}

void orc_cot_receive(orc_label *result, const uint8_t *flags, orc_label seed,
                     const orc_label *msgs, uint64_t n) {
    orc_mitccrh m;
    orc_mitccrh_init(&m, seed, OT_BATCH);
    orc_label pad[OT_BATCH];
    memset(pad, 0, sizeof pad);
    for (uint64_t i = 0; i < n; i += OT_BATCH) {
        uint64_t end = OT_BATCH;
        if (end > n - i) end = n - i;
        for (uint64_t j = 0; j < end; j++) pad[j] = result[i + j];     /* copy(pad, result[i:]) */
        orc_mitccrh_hash(&m, pad, OT_BATCH, 1);
        for (uint64_t j = 0; j < end; j++) {
            orc_label r = flags[i + j] ? msgs[2 * (i + j) + 1] : msgs[2 * (i + j)];
            result[i + j] = lxor(r, pad[j]);
        }
    }
}

void orc_rot_send(const orc_label *data, orc_label delta, orc_label seed, orc_wire *wires, uint64_t n) {
    orc_mitccrh m;
    orc_mitccrh_init(&m, seed, OT_BATCH);
    orc_label pad[2 * OT_BATCH];
    memset(pad, 0, sizeof pad);
    for (uint64_t i = 0; i < n; i += OT_BATCH) {
        uint64_t end = i + OT_BATCH;
        if (end > n) end = n;
        for (uint64_t j = i; j < end; j++) {
            pad[2 * (j - i)] = data[j];
            pad[2 * (j - i) + 1] = lxor(data[j], delta);
        }
        orc_mitccrh_hash(&m, pad, OT_BATCH, 2);
        for (uint64_t j = i; j < end; j++) {
            wires[j].l0 = pad[2 * (j - i)];
            wires[j].l1 = pad[2 * (j - i) + 1];
        }
    }
}

void orc_rot_receive(orc_label *result, orc_label seed, uint64_t n) {
    orc_mitccrh m;
    orc_mitccrh_init(&m, seed, OT_BATCH);
    orc_label pad[OT_BATCH];
    memset(pad, 0, sizeof pad);
    for (uint64_t i = 0; i < n; i += OT_BATCH) {
        uint64_t cnt = OT_BATCH;
        if (cnt > n - i) cnt = n - i;
        for (uint64_t j = 0; j < cnt; j++) pad[j] = result[i + j];
        orc_mitccrh_hash(&m, pad, OT_BATCH, 1);
        for (uint64_t j = 0; j < cnt; j++) result[i + j] = pad[j];
    }
}

/* ======================================================================== */
/* GF(2^128) carry-less multiply without reduction                             */
/* ot/mul128_generic.go:9-47 (clmul64 shift-and-xor), gf128.go:14-27           */
/* Polynomial coefficient i is Label.Bit(i): D0 holds x^0..x^63.               */
/* ======================================================================== */
static void clmul64(uint64_t a, uint64_t b, uint64_t *lo, uint64_t *hi) {
    uint64_t l = 0, h = 0;
    for (int i = 0; i < 64; i++)
        if ((b >> i) & 1) {
            if (i == 0) l ^= a;
            else { l ^= a << i; h ^= a >> (64 - i); }
        }
    *lo = l; *hi = h;
}
void orc_mul128(orc_label a, orc_label b, orc_label *lo, orc_label *hi) {
    uint64_t p00l, p00h, p01l, p01h, p10l, p10h, p11l, p11h;
    clmul64(a.d0, b.d0, &p00l, &p00h);
    clmul64(a.d0, b.d1, &p01l, &p01h);
    clmul64(a.d1, b.d0, &p10l, &p10h);
    clmul64(a.d1, b.d1, &p11l, &p11h);
    uint64_t midl = p01l ^ p10l, midh = p01h ^ p10h;
    lo->d0 = p00l; lo->d1 = p00h ^ midl;
    hi->d0 = midh ^ p11l; hi->d1 = p11h;
}
void orc_inner_product(const orc_label *a, const orc_label *b, uint64_t n, orc_label *lo, orc_label *hi) {
    orc_label r1 = {0, 0}, r2 = {0, 0};
    for (uint64_t i = 0; i < n; i++) {
        orc_label l, h;
        orc_mul128(a[i], b[i], &l, &h);
        r1 = lxor(r1, l); r2 = lxor(r2, h);
    }
    *lo = r1; *hi = r2;
}

/* ======================================================================== */
/* Malicious-mode consistency check of IKNP (ot/iknp.go:137-192 sender,       */
/* :373-465 receiver): chi_i = prgLabels(newPrg(seed2)) in stream order, i.e.  */
/* label number chi_start + i is keystream block chi_start + i read with       */
/* SetBytes (big-endian); sums over `labels`: (lo, hi) ^= mul128(chi_i, l_i),  */
/* and for the receiver x ^= chi_i where choice[i] is set (the And(select1) /  */
/* And(select0) of :425-431).  out = {lo, hi, x}.  The 1024-label chunking of  */
/* the reference is only buffer management: the stream is continuous.          */
/* ======================================================================== */
void orc_iknp_check_sums(orc_label seed2, uint64_t chi_start, const orc_label *labels,
                         const uint8_t *choice, uint64_t n, orc_label out[3]) {
    orc_label lo = {0, 0}, hi = {0, 0}, x = {0, 0};
    for (uint64_t i = 0; i < n; i++) {
        uint8_t buf[16];
        orc_prg(seed2, (chi_start + i) * 16, buf, 16);
        orc_label chi = {get_be64(buf), get_be64(buf + 8)};          /* SetBytes, label.go:111-114 */
        orc_label l, h;
        orc_mul128(chi, labels[i], &l, &h);
        lo = lxor(lo, l); hi = lxor(hi, h);
        if (choice && choice[i]) x = lxor(x, chi);
    }
    out[0] = lo; out[1] = hi; out[2] = x;
}
