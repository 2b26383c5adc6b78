/*
 * gcb_oracle.h -- CPU restatement of the reference's garble / eval / streaming /
 * IKNP / MiTCCRH algorithms.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under mpc_b200/ (the product) includes,
 * links or calls this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, as the checker or as the CPU
 * baseline -- never as the thing shipped.
 *
 * Each function cites the reference file:line it restates (paths relative to
 * the markkurossi/mpc tree).  The reference's arithmetic that lives in the Go
 * standard library (crypto/aes, crypto/cipher.NewCTR) is restated from the
 * public standards FIPS-197 and SP 800-38A and pinned by their vectors.
 *
 * PARITY PINNING STATUS (see DESIGN.md section "Oracle"):
 *   pinned by the reference's own vectors: MiTCCRH (ot/mitccrh_test.go:22-31),
 *   label algebra (ot/label_test.go:38-91), Gate layout
 *   (circuit/circuit_test.go:14-19), functional digest of sha2pc
 *   (sha2pc/sha2pc_test.go:124) and the plaintext KATs of the shipped
 *   circuits; AES / CTR by FIPS-197 App. C and SP 800-38A F.5.
 *   Circuit.Garble's bytes (garbled rows in gate order, input and output wire
 *   labels; 32-byte key) are pinned by the reference's golden transcript
 *   hashes (sha2pc/sha2pc_test.go:74-131, TestDeterministicTranscript): the
 *   test's math/rand streams, crypto/rand.Int, P-256 and the round encodings
 *   are restated in tests/gostd.py + tests/sha2pc_transcript.py, and
 *   tests/test_reference_transcript.py reproduces all four hashes with
 *   orc_garble / orc_eval in the middle (and with the CUDA path on the GPU).
 *   PARITY UNPINNED at the byte level for IKNP label bytes and for the
 *   Streaming.Garble record stream: the reference holds no vector for them.
 *   The stream shares garbleGate's arithmetic with the pinned Garble; IKNP
 *   rests on its pinned parts (AES-CTR by SP 800-38A, MiTCCRH golden blocks),
 *   an independent Python restatement on OpenSSL's AES (tests/pyref.py) and
 *   the protocol invariant t = q ^ b * Delta.
 */
#ifndef GCB_ORACLE_H
#define GCB_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ot.Label in Go memory order (ot/label.go:28-31): D0 is the HIGH half. */
typedef struct { uint64_t d0, d1; } orc_label;
/* ot.Wire (ot/label.go:18-21). */
typedef struct { orc_label l0, l1; } orc_wire;
/* circuit.Gate, 20 bytes (circuit/circuit.go:260-266). */
typedef struct {
    uint32_t in0, in1, out;
    uint8_t op;
    uint8_t pad[3];
    uint32_t level;
} orc_gate;

enum { ORC_XOR = 0, ORC_XNOR = 1, ORC_AND = 2, ORC_OR = 3, ORC_INV = 4 };

/* error codes (negative), mirroring the reference's error returns */
enum {
    ORC_OK = 0,
    ORC_E_KEYLEN = -1,        /* aes.NewCipher: invalid key size */
    ORC_E_BADOP = -2,         /* "invalid gate type" / "invalid operation" */
    ORC_E_AND_ROWS = -3,      /* "corrupted ciruit: AND row length" (eval.go:55) */
    ORC_E_ROW_INDEX = -4,     /* "corrupted circuit: index >= row" (eval.go:87,103) */
    ORC_E_BUFFER = -5,        /* output buffer too small */
    ORC_E_CHUNK = -6,         /* "invalid chunk size" (iknp.go:207) */
    ORC_E_ARG = -7
};

/* ---- AES (Go stdlib crypto/aes == FIPS-197) ---------------------------- */
int  orc_have_aesni(void);
void orc_set_aesni(int on);       /* 0 = force the portable byte-wise path */
int  orc_aes_encrypt_block(const uint8_t *key, uint32_t keylen,
                           const uint8_t in[16], uint8_t out[16]);

/* ---- label algebra (ot/label.go:57-96) and gate hashes ------------------ */
orc_label orc_label_mul2(orc_label l);
orc_label orc_label_mul4(orc_label l);
void orc_label_to_bytes(orc_label l, uint8_t out[16]);   /* GetData :105-108 */
orc_label orc_label_from_bytes(const uint8_t in[16]);    /* SetData :111-114 */
/* circuit/garble.go:104-136 */
int orc_encrypt_half(const uint8_t *key, uint32_t keylen, orc_label x, uint32_t tweak, orc_label *out);
/* circuit/garble.go:40-56 and :58-73 */
int orc_encrypt(const uint8_t *key, uint32_t keylen, orc_label a, orc_label b, orc_label c,
                uint32_t t, orc_label *out);
int orc_decrypt(const uint8_t *key, uint32_t keylen, orc_label a, orc_label b, uint32_t t,
                orc_label c, orc_label *out);

/* ---- Circuit.Garble / Circuit.Eval -------------------------------------- */
/* circuit/garble.go:248-308.  `rand` is the byte stream the Go code would read
 * from its io.Reader: 16 bytes for R, then 16 per input wire.  row_off (may be
 * NULL) receives ngates+1 slab offsets: Gates[i] == slab[row_off[i]:row_off[i+1]]. */
int orc_garble(const orc_gate *gates, uint32_t ngates, uint32_t nwires, uint32_t ninputs,
               const uint8_t *key, uint32_t keylen, const uint8_t *rand,
               orc_label *r_out, orc_wire *wires, orc_label *slab, uint32_t *row_off);
/* circuit/eval.go:17-115.  wires[0..ninputs) pre-filled, all nwires written. */
int orc_eval(const orc_gate *gates, uint32_t ngates, uint32_t nwires,
             const uint8_t *key, uint32_t keylen, orc_label *wires,
             const orc_label *slab, const uint32_t *row_off);
/* Batch drivers used as the CPU baseline: instances split over `threads`
 * pthreads, each instance exactly one orc_garble / orc_eval call.
 * Layouts: rand [batch][16*(1+ninputs)], keys [batch][keylen] if key_stride
 * != 0 else one shared key, tables [batch][rows], io_wires
 * [batch][ninputs+noutputs] (input wires then output wires). */
int orc_garble_batch(const orc_gate *gates, uint32_t ngates, uint32_t nwires, uint32_t ninputs,
                     uint32_t noutputs, const uint8_t *keys, uint32_t keylen, uint32_t key_stride,
                     uint32_t batch, const uint8_t *rand, orc_label *r_out, orc_label *tables,
                     orc_wire *io_wires, int threads);
int orc_eval_batch(const orc_gate *gates, uint32_t ngates, uint32_t nwires, uint32_t ninputs,
                   uint32_t noutputs, const uint8_t *keys, uint32_t keylen, uint32_t key_stride,
                   uint32_t batch, const orc_label *tables, const orc_label *in_labels,
                   orc_label *out_labels, int threads);

/* ---- Streaming garbler (circuit/stream_garble.go) ----------------------- */
typedef struct orc_stream orc_stream;
/* NewStreaming :41-76: rand = 16 bytes R then 16 per input id (in slice order). */
orc_stream *orc_stream_new(const uint8_t *key, uint32_t keylen, const uint8_t *rand,
                           const uint32_t *input_ids, uint32_t ninputs);
void orc_stream_free(orc_stream *s);
orc_label orc_stream_r(const orc_stream *s);
orc_wire orc_stream_get_input(const orc_stream *s, uint32_t id);          /* :117-119 */
void orc_stream_set_wire(orc_stream *s, uint32_t id, orc_wire w);
/* Streaming.Garble :161-191 + garbleGate :195-449: emits the gate records. */
int orc_stream_garble(orc_stream *s, const orc_gate *gates, uint32_t ngates, uint32_t nwires,
                      const uint32_t *in, uint32_t nin, const uint32_t *out, uint32_t nout,
                      uint8_t *buf, size_t cap, size_t *written);

/* ---- Streaming evaluator (circuit/stream_evaluator.go:29-96,270-432) ---- */
typedef struct orc_seval orc_seval;
orc_seval *orc_seval_new(const uint8_t *key, uint32_t keylen);
void orc_seval_free(orc_seval *s);
void orc_seval_set(orc_seval *s, uint32_t id, orc_label l);
orc_label orc_seval_get(const orc_seval *s, uint32_t id);
int orc_seval_circuit(orc_seval *s, const uint8_t *buf, size_t len, uint32_t ngates,
                      uint32_t ntmp, uint32_t nwires, size_t *consumed);

/* ---- IKNP (ot/iknp.go) --------------------------------------------------- */
/* newPrg/prg :622-637: AES-128 keyed by BE(label), CTR, zero IV, stateful and
 * byte granular; `pos` is the number of keystream bytes already consumed. */
void orc_prg(orc_label key, uint64_t pos, uint8_t *buf, size_t n);
/* createLabels :647-683 */
void orc_create_labels(orc_label *l, size_t nl, const uint8_t *buf, int w);
/* IKNPReceiver.receive :468-511.  choice: n bytes of 0/1.  u_out receives the
 * concatenation of the chunks passed to SendData ((sum byteRows)*128 bytes);
 * *pos is the shared keystream position of all 256 PRGs, advanced on return. */
int orc_iknp_receive(const orc_label k0[128], const orc_label k1[128], uint64_t *pos,
                     const uint8_t *choice, uint64_t n, uint8_t *u_out, size_t u_cap,
                     size_t *u_len, orc_label *result);
/* IKNPSender.send :197-226 on the same chunk stream. */
int orc_iknp_send(const orc_label k[128], orc_label delta, uint64_t *pos, const uint8_t *u,
                  size_t u_len, uint64_t n, orc_label *result);
/* Bit-COT variants :259-310, :554-620 (packed LSB-first uint64 words). */
int orc_iknp_receive_bits(const orc_label k0[128], const orc_label k1[128], uint64_t *pos,
                          const uint64_t *choices, uint64_t n, uint8_t *u_out, size_t u_cap,
                          size_t *u_len, uint64_t *result);
int orc_iknp_send_bits(const orc_label k[128], orc_label delta, uint64_t *pos, const uint8_t *u,
                       size_t u_len, uint64_t n, uint64_t *result);

/* ---- MiTCCRH (ot/mitccrh.go) ---------------------------------------------- */
typedef struct {
    int batch_size;
    orc_label start;
    uint64_t gid;
    int key_used;
    uint8_t keys[64][16];
} orc_mitccrh;
void orc_mitccrh_init(orc_mitccrh *m, orc_label seed, int batch_size);       /* :61-68 */
int  orc_mitccrh_hash(orc_mitccrh *m, orc_label *blks, int k, int h);        /* :93-128 */

/* ---- COT / ROT post-processing (ot/cot.go:136-235, ot/rot.go:132-202) ---- */
/* Sender: data = IKNP q labels; out_msgs receives the 2n labels put on the
 * wire (COT) or wires receives the random wire pairs (ROT). */
void orc_cot_send(const orc_label *data, orc_label delta, orc_label seed, const orc_wire *wires,
                  uint64_t n, orc_label *out_msgs);
void orc_cot_receive(orc_label *result, const uint8_t *flags, orc_label seed,
                     const orc_label *msgs, uint64_t n);
void orc_rot_send(const orc_label *data, orc_label delta, orc_label seed, orc_wire *wires, uint64_t n);
void orc_rot_receive(orc_label *result, orc_label seed, uint64_t n);

/* ---- GF(2^128) carry-less multiply (ot/mul128_ref.go, gf128.go:14) -------- */
void orc_mul128(orc_label a, orc_label b, orc_label *lo, orc_label *hi);
void orc_inner_product(const orc_label *a, const orc_label *b, uint64_t n, orc_label *lo, orc_label *hi);
/* Sums of the malicious-mode check (ot/iknp.go:150-173, :408-451): out = {lo, hi, x}. */
void orc_iknp_check_sums(orc_label seed2, uint64_t chi_start, const orc_label *labels,
                         const uint8_t *choice, uint64_t n, orc_label out[3]);

#ifdef __cplusplus
}
#endif
#endif
