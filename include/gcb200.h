/*
 * gcb200.h -- C ABI of the B200 garbled-circuit / OT-extension engine.
 *
 * This is the drop-in boundary: exactly the entry points a cgo shim behind the
 * reference's unchanged Go API (packages circuit and ot of markkurossi/mpc)
 * binds.  Each entry point cites the reference interface it replaces; the Go
 * bindings are shown in INTEGRATION.md and kept under go/.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer of the host entry points is a
 *     HOST pointer valid for the duration of the call (cgo rule: nothing is
 *     retained).  The *_dev entry points take DEVICE pointers and a CUDA stream
 *     (cudaStream_t passed as void*) and only enqueue work.
 *   - labels are 16 bytes in Go memory order: {uint64 D0; uint64 D1}, D0 the
 *     high half (ot/label.go:28-31); wires are {L0, L1} (ot/label.go:18-21);
 *     gates are the 20-byte circuit.Gate (circuit/circuit.go:260-266).
 *   - every function returns 0 on success or a negative gcb_status;
 *     gcb_last_error() gives the thread-local message.
 *   - all entry points are re-entrant; plans are immutable after creation and
 *     may be shared by concurrent callers (circuit.Circuit is shared by
 *     goroutines, sha2pc/sha256xor_circuit.go:16-23).
 *   - there is no CPU fallback: without a CUDA device every compute entry point
 *     fails with GCB_E_CUDA.
 */
#ifndef GCB200_H
#define GCB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t d0, d1; } gcb_label;          /* ot.Label */
typedef struct { gcb_label l0, l1; } gcb_wire;          /* ot.Wire  */
typedef struct {                                        /* circuit.Gate, 20 bytes */
    uint32_t in0, in1, out;
    uint8_t op;                                         /* 0 XOR 1 XNOR 2 AND 3 OR 4 INV */
    uint8_t pad[3];
    uint32_t level;
} gcb_gate;

typedef enum {
    GCB_OK = 0,
    GCB_E_ARG = -1,          /* bad argument (null pointer, zero size, ...) */
    GCB_E_KEYLEN = -2,       /* aes.NewCipher: key must be 16, 24 or 32 bytes */
    GCB_E_BADOP = -3,        /* "invalid gate type" (circuit/garble.go:325) */
    GCB_E_WIRE = -4,         /* wire index out of range / read before assignment */
    GCB_E_CUDA = -5,         /* no device, launch or memory failure */
    GCB_E_TOO_LARGE = -6,    /* live wire set does not fit on chip */
    GCB_E_BUFFER = -7,       /* output buffer too small */
    GCB_E_CHUNK = -8,        /* "invalid chunk size" (ot/iknp.go:207) */
    GCB_E_CORRUPT = -9       /* "corrupted circuit" (circuit/eval.go:55,87,103) */
} gcb_status;

const char *gcb_last_error(void);
const char *gcb_version(void);
/* Diagnostic: kernels this library has launched in this process so far (bench.py reports the difference
 * over its timed region as `gpu_launches`). */
uint64_t gcb_launch_count(void);

/* Device selection.  cgo calls may arrive on any OS thread, so every entry point
 * re-selects its device.
 *   gcb_set_devices(ids, n)  the process-wide engine setting of SURVEY.md section 8(b): the devices the
 *                            library may use (n = 0 clears it).  Entry points whose work splits into
 *                            independent units -- host-pointer gcb_garble / gcb_eval (and _begin) by
 *                            contiguous blocks of instances, gcb_iknp_*_expand by 512-row chunk ranges,
 *                            gcb_mitccrh_hash / gcb_cot_* / gcb_rot_* by OT ranges, and the _dev forms of
 *                            garble / eval with GCB_FLAG_FANOUT -- fan one call out over all of them
 *                            (circuit/garble.go:285-299 is the loop being parallelised); everything else
 *                            runs on ids[0].  Results are byte-identical to a one-device call.
 *   gcb_set_device(d)        the calling thread's override: d >= 0 pins this thread's calls to one
 *                            device, -1 returns it to the process-wide setting.
 * Default with neither set: device LOCAL_RANK if that variable is set, else device 0.
 * gcb_get_devices: the devices a call of this thread would use (returns their number). */
int gcb_set_device(int device);
int gcb_set_devices(const int *ids, int n);
int gcb_get_devices(int *ids, int cap);
int gcb_device_count(void);

/* Page-locked host memory.  Replaces the per-circuit scratch pool of
 * Circuit.Garble (garbleScratchPool, circuit/garble.go:193-224; Garbled.Release
 * :232-244 returns to it): slabs allocated here are DMA'd in place by the host
 * entry points; any other host memory is staged through an internal pinned
 * buffer with one extra copy.  Returns NULL on failure. */
void *gcb_host_alloc(size_t bytes);
void gcb_host_free(void *p);

/* Device buffers and streams for callers without a CUDA binding of their own (Go): what the _dev
 * entry points take.  Keeping the operands of consecutive calls on the device -- IKNP expansion ->
 * COT post-processing, garble -> wire format -> eval -- removes the PCIe round trips between them.
 * A stream handle is a cudaStream_t (NULL = the default stream); copies are asynchronous on the
 * stream when the host side is page-locked (gcb_host_alloc), and gcb_dev_sync waits for it. */
void *gcb_dev_alloc(size_t bytes);
void gcb_dev_free(void *p);
int gcb_dev_upload(void *dst_dev, const void *src_host, size_t bytes, void *stream);
int gcb_dev_download(void *dst_host, const void *src_dev, size_t bytes, void *stream);
int gcb_dev_stream_create(void **stream);
void gcb_dev_stream_destroy(void *stream);
int gcb_dev_sync(void *stream);

/* ------------------------------------------------------------------ plans --- */
/* A plan is the compiled, immutable form of one circuit.Circuit: gates grouped
 * into dependency steps, with the static per-gate tweak ids (the `id` counter of
 * circuit/garble.go:357-359,419-420,451-452), slab row offsets
 * (circuit/garble.go:292-298) and on-chip wire slots assigned from liveness. */
typedef struct gcb_plan gcb_plan;

typedef struct {
    uint32_t num_gates, num_wires, num_inputs, num_outputs;
    uint32_t num_rows;        /* garbled-table labels per instance (slab size) */
    uint32_t num_tweaks;
    uint32_t num_steps;       /* dependency steps (barriers per instance) */
    uint32_t num_slots;       /* peak live wire labels held on chip per instance */
    uint32_t num_and, num_or, num_inv, num_free;
    uint32_t teams_per_sm;    /* instances resident per SM with this plan */
    uint32_t team_threads;
    uint32_t garble_hashes;   /* AES blocks per garbled instance (4/AND, 4/OR, 2/INV) */
    uint32_t eval_hashes;     /* AES blocks per evaluated instance (2/AND, 1/OR, 1/INV) */
    uint32_t garble_passes;   /* warp passes of 32 AES blocks the cipher levels start, garbler ... */
    uint32_t eval_passes;     /* ... and evaluator (>= hashes / 32: a level pays for whole passes) */
    uint32_t num_hot_slots;   /* of num_slots, the labels held in shared memory; the rest (values that idle for
                                 hundreds of levels between uses) live in an L2-resident scratch per instance */
} gcb_plan_info;

/* Replaces: the per-circuit preparation Circuit.Garble does lazily
 * (garbleScratchPool, circuit/garble.go:193-224).  Input is Circuit.Gates as is. */
int gcb_plan_create(const gcb_gate *gates, uint32_t num_gates, uint32_t num_wires,
                    uint32_t num_inputs, uint32_t num_outputs, gcb_plan **out);
void gcb_plan_destroy(gcb_plan *plan);
int gcb_plan_get_info(const gcb_plan *plan, gcb_plan_info *info);
/* The same for the plan a call with `batch` instances per device runs on: deep, narrow circuits whose batch
 * overflows the resident instances of the default plan get a second plan with split live ranges (more instances
 * per SM, num_hot_slots < num_slots), built on first use.  The output bytes never depend on the plan.
 * Replaces nothing in the reference (diagnostics: bench.py names the kernel geometry it timed). */
int gcb_plan_get_info_for_batch(const gcb_plan *plan, uint64_t batch, gcb_plan_info *info);
/* Static slab offset of every gate's first row (num_gates+1 entries): lets the
 * Go side rebuild Garbled.Gates [][]ot.Label slice headers over the slab. */
int gcb_plan_row_offsets(const gcb_plan *plan, uint32_t *row_off);

/* ----------------------------------------------------------- garble / eval --- */
#define GCB_FLAG_NONE 0u
/* gcb_garble_dev / gcb_eval_dev only: the operands live on the calling thread's device; the batch is
 * split over the gcb_set_devices list, inputs scattered and results gathered with peer copies over
 * NVLink, sliced behind the kernels and ordered against `stream` (no wires_full). */
#define GCB_FLAG_FANOUT 1u

/* Replaces (*Circuit).Garble(rand, key) (circuit/garble.go:248-308), batched.
 *   keys       : key bytes; key_stride 0 = one key shared by the batch, else
 *                instance i uses keys + i*key_stride.  keylen 16/24/32.
 *   r          : [batch] the 16 bytes read for R per instance, as a label;
 *                the S bit is forced to 1 inside (garble.go:258).
 *   in_l0      : [batch][num_inputs] the L0 label read for each input wire
 *                (garble.go:271-278); L1 = L0 ^ R is derived.
 *   tables     : [batch][num_rows] out, the slab: rows of gate g at
 *                row_off[g] in original gate order.
 *   io_wires   : [batch][num_inputs+num_outputs] out (may be NULL): Wires[0..in)
 *                then Wires[NumWires-out..NumWires) -- what the callers read
 *                (circuit/garbler.go:85-102,153; sha2pc/garbler.go:113-126).
 *   wires_full : [batch][num_wires] out (may be NULL): the complete Garbled.Wires.
 * batch = 1 reproduces one Go call exactly. */
int gcb_garble(const gcb_plan *plan, const uint8_t *keys, uint32_t keylen, uint32_t key_stride,
               uint32_t batch, const gcb_label *r, const gcb_label *in_l0, gcb_label *tables,
               gcb_wire *io_wires, gcb_wire *wires_full, uint32_t flags);

/* Replaces (*Circuit).Eval(key, wires, garbled) (circuit/eval.go:17-115), batched.
 *   tables     : [batch][num_rows] the slab (row counts per gate are static, so
 *                the "corrupted circuit" length checks are done by the caller
 *                against gcb_plan_row_offsets before the call).
 *   in_labels  : [batch][num_inputs]   wires[0..Inputs.Size())
 *   out_labels : [batch][num_outputs]  wires[NumWires-Outputs.Size()..)
 *   wires_full : [batch][num_wires] out (may be NULL): every wire label. */
int gcb_eval(const gcb_plan *plan, const uint8_t *keys, uint32_t keylen, uint32_t key_stride,
             uint32_t batch, const gcb_label *tables, const gcb_label *in_labels,
             gcb_label *out_labels, gcb_label *wires_full, uint32_t flags);

/* The same calls split in two, so that ONE host thread keeps the copy engines and the SMs of every selected
 * device busy: _begin queues the whole call -- input copies, kernels slice by slice, result copies -- and
 * returns; gcb_job_wait blocks until the results are in the caller's buffers, frees the job and returns its
 * status (gcb_job_done polls: 1 = finished).  Every buffer must stay valid and untouched until the wait; Go
 * callers pass gcb_host_alloc memory (C memory: the cgo pointer rules do not apply to it), which is also
 * what makes the copies truly asynchronous -- pageable memory is staged with a host memcpy inside _begin
 * (inputs) and inside the wait (results).  Many jobs may be in flight; they complete in any order. */
typedef struct gcb_job gcb_job;
int gcb_garble_begin(const gcb_plan *plan, const uint8_t *keys, uint32_t keylen, uint32_t key_stride,
                     uint32_t batch, const gcb_label *r, const gcb_label *in_l0, gcb_label *tables,
                     gcb_wire *io_wires, gcb_wire *wires_full, uint32_t flags, gcb_job **job);
int gcb_eval_begin(const gcb_plan *plan, const uint8_t *keys, uint32_t keylen, uint32_t key_stride,
                   uint32_t batch, const gcb_label *tables, const gcb_label *in_labels,
                   gcb_label *out_labels, gcb_label *wires_full, uint32_t flags, gcb_job **job);
int gcb_job_wait(gcb_job *job);
int gcb_job_done(gcb_job *job);

/* Device-resident variants (all pointers are device pointers; asynchronous on
 * `stream`).  Used for throughput measurement and multi-stage pipelines. */
int gcb_garble_dev(const gcb_plan *plan, const uint8_t *keys, uint32_t keylen, uint32_t key_stride,
                   uint32_t batch, const gcb_label *r, const gcb_label *in_l0, gcb_label *tables,
                   gcb_wire *io_wires, gcb_wire *wires_full, uint32_t flags, void *stream);
int gcb_eval_dev(const gcb_plan *plan, const uint8_t *keys, uint32_t keylen, uint32_t key_stride,
                 uint32_t batch, const gcb_label *tables, const gcb_label *in_labels,
                 gcb_label *out_labels, gcb_label *wires_full, uint32_t flags, void *stream);

/* Input selection and output decoding on the device, the label plumbing either
 * side of Garble/Eval: LabelForBit / BitFromLabel (circuit/helpers.go:10-27).
 *   bits: one byte (0/1) per wire.  gcb_decode_bits_dev writes 0/1, or 2 for
 *   "unknown label". */
int gcb_select_labels_dev(const gcb_wire *wires, size_t wire_stride, const uint8_t *bits,
                          gcb_label *out, uint32_t batch, uint32_t n, void *stream);
int gcb_decode_bits_dev(const gcb_wire *wires, size_t wire_stride, const gcb_label *labels,
                        uint8_t *bits, uint32_t batch, uint32_t n, void *stream);

/* ------------------------------------------------------------- gate hashes --- */
/* Replaces the hash micro-benchmarks: encryptHalf (circuit/garble.go:104-136,
 * BenchmarkEncHalf circuit/enc_test.go:73) and the stand-alone AES-NI benchmark
 * circuit/aesni/c/aesni.c.  out[i] = H(x[i], tweak0 + i) = AES_k(K) ^ K,
 * K = 2*x[i] ^ (tweak0+i). */
int gcb_hash_half(const uint8_t *key, uint32_t keylen, const gcb_label *x, uint32_t tweak0,
                  gcb_label *out, uint64_t n);
int gcb_hash_half_dev(const uint8_t *key, uint32_t keylen, const gcb_label *x, uint32_t tweak0,
                      gcb_label *out, uint64_t n, void *stream);

/* --------------------------------------------------------------- streaming --- */
/* Replaces circuit.Streaming (circuit/stream_garble.go): NewStreaming, Get/Set,
 * GetInput(s) and Streaming.Garble, for `batch` independent program instances
 * garbled in lock step (batch = 1 is the Go object). */
typedef struct gcb_stream gcb_stream;

/* NewStreaming (stream_garble.go:41-76): r [batch] raw R labels (S forced),
 * input_ids [ninputs] permanent wire ids, in_l0 [batch][ninputs]. */
int gcb_stream_create(const uint8_t *keys, uint32_t keylen, uint32_t key_stride, uint32_t batch,
                      const gcb_label *r, const uint32_t *input_ids, uint32_t ninputs,
                      const gcb_label *in_l0, gcb_stream **out);
void gcb_stream_destroy(gcb_stream *s);
/* GetInput / GetInputs (stream_garble.go:117-128): wires [batch][n]. */
int gcb_stream_get_wires(gcb_stream *s, const uint32_t *ids, uint32_t n, gcb_wire *wires);
/* Size in bytes of the record stream Streaming.Garble emits for this step. */
int gcb_stream_step_size(gcb_stream *s, const gcb_plan *plan, const uint32_t *in, uint32_t nin,
                         const uint32_t *out, uint32_t nout, size_t *bytes);
/* Streaming.Garble (stream_garble.go:161-191) + garbleGate (:195-449): garbles
 * the sub-circuit for every instance and writes, per instance, the exact byte
 * stream the reference puts into conn.WriteBuf (gate records: op|flags byte,
 * BE u16/u32 wire ids, BE rows).  dst: [batch][dst_stride] host bytes.
 * ns_init / ns_garble mirror the two durations the Go method returns. */
int gcb_stream_garble(gcb_stream *s, const gcb_plan *plan, const uint32_t *in, uint32_t nin,
                      const uint32_t *out, uint32_t nout, uint8_t *dst, size_t dst_stride,
                      size_t *written, uint64_t *ns_init, uint64_t *ns_garble);
/* The same, split in two so that a caller can keep one step in flight: _begin queues the gate kernel, the
 * serialiser and the copy into `dst` and returns; _wait(s, 1) returns when the bytes of every step but the
 * one begun last are in place, _wait(s, 0) when all are.  The gate kernel of step k+1 then runs while the
 * record bytes of step k are still crossing PCIe (the wire file is on the device, so step k+1 does not
 * depend on that copy).  `dst` of a step must stay untouched until a _wait has covered it; a third _begin
 * first waits for the first step's copy (two staging sets). */
int gcb_stream_garble_begin(gcb_stream *s, const gcb_plan *plan, const uint32_t *in, uint32_t nin,
                            const uint32_t *out, uint32_t nout, uint8_t *dst, size_t dst_stride, size_t *written);
int gcb_stream_garble_wait(gcb_stream *s, uint32_t leave_in_flight);

/* ------------------------------------------------------ streaming evaluator --- */
/* Replaces circuit.StreamEval (circuit/stream_evaluator.go:29-96: NewStreamEval,
 * Get, Set, SetInputs, InitCircuit) and the gate loop of StreamEvaluator for one
 * OpCircuit body (:270-432), for `batch` program instances in lock step. */
typedef struct gcb_seval gcb_seval;
int gcb_seval_create(const uint8_t *keys, uint32_t keylen, uint32_t key_stride, uint32_t batch,
                     gcb_seval **out);
void gcb_seval_destroy(gcb_seval *s);
/* labels: [batch][n] */
int gcb_seval_set_wires(gcb_seval *s, const uint32_t *ids, uint32_t n, const gcb_label *labels);
int gcb_seval_get_wires(gcb_seval *s, const uint32_t *ids, uint32_t n, gcb_label *labels);
/* src: [batch][src_stride] bytes received after the OpCircuit header (step,
 * numGates, numTmpWires, numWires: compiler/ssa/streamer.go:679-693): `ngates`
 * gate records; *consumed = bytes of one instance's records.  All instances
 * must carry the same gate headers (same circuit and wire ids).  Malformed
 * records give GCB_E_BUFFER / GCB_E_BADOP / GCB_E_CORRUPT. */
int gcb_seval_circuit(gcb_seval *s, const uint8_t *src, size_t src_stride, size_t len, uint32_t ngates,
                      uint32_t ntmp, uint32_t nwires, size_t *consumed);

/* ------------------------------------------------------------------- IKNP --- */
/* Replaces the inner loops of IKNPReceiver.receive (ot/iknp.go:468-511) and
 * IKNPSender.send (ot/iknp.go:197-226): AES-128-CTR column expansion
 * (newPrg/prg :622-637), the U = T0 ^ T1 ^ b matrix, and the 128-wide bit
 * transpose createLabels (:647-683).  Base OTs, framing and I/O stay in Go;
 * `stream_pos` is the number of keystream bytes every column PRG has already
 * produced (the Go streams are stateful and byte granular); the new position
 * is stream_pos + gcb_iknp_stream_advance(n).
 *   u layout: the concatenation of the chunks the receiver passes to SendData:
 *   chunk c covers rows [512c, 512c+rows), byteRows = ceil(rows/8), column i at
 *   u[chunk_off + i*byteRows ...]; total gcb_iknp_u_size(n) bytes.
 *   choice: n bytes of 0/1.  labels: [n] in Go memory order. */
size_t gcb_iknp_u_size(uint64_t n);
uint64_t gcb_iknp_stream_advance(uint64_t n);
int gcb_iknp_receiver_expand(const gcb_label k0[128], const gcb_label k1[128], uint64_t stream_pos,
                             const uint8_t *choice, uint64_t n, uint8_t *u_out, gcb_label *labels);
int gcb_iknp_sender_expand(const gcb_label k[128], const gcb_label *delta, uint64_t stream_pos,
                           const uint8_t *u, size_t u_len, uint64_t n, gcb_label *labels);
int gcb_iknp_receiver_expand_dev(const gcb_label *k0, const gcb_label *k1, uint64_t stream_pos,
                                 const uint8_t *choice, uint64_t n, uint8_t *u_out,
                                 gcb_label *labels, void *stream);
int gcb_iknp_sender_expand_dev(const gcb_label *k, const gcb_label *delta, uint64_t stream_pos,
                               const uint8_t *u, size_t u_len, uint64_t n, gcb_label *labels,
                               void *stream);

/* Bit-COT variants.  Replace IKNPReceiver.ReceiveBits (ot/iknp.go:554-620) and
 * IKNPSender.SendBits (ot/iknp.go:259-310), used by gmw/triples.go:335-422: the
 * same U exchange, the output is Bit(0) of every label packed LSB-first into
 * uint64 words ((n+63)/64 words, overwritten).  choices: the packed []uint64 of
 * the Go API; as in the reference they are XORed into U in whole 64-bit words
 * (words = byteRows / 8 per chunk), and the _dev variant reads choice words up to
 * the end of the last 512-row chunk. */
int gcb_iknp_receiver_expand_bits(const gcb_label k0[128], const gcb_label k1[128], uint64_t stream_pos,
                                  const uint64_t *choices, uint64_t n, uint8_t *u_out, uint64_t *result);
int gcb_iknp_sender_expand_bits(const gcb_label k[128], const gcb_label *delta, uint64_t stream_pos,
                                const uint8_t *u, size_t u_len, uint64_t n, uint64_t *result);
int gcb_iknp_receiver_expand_bits_dev(const gcb_label *k0, const gcb_label *k1, uint64_t stream_pos,
                                      const uint64_t *choices, uint64_t n, uint8_t *u_out, uint64_t *result,
                                      void *stream);
int gcb_iknp_sender_expand_bits_dev(const gcb_label *k, const gcb_label *delta, uint64_t stream_pos,
                                    const uint8_t *u, size_t u_len, uint64_t n, uint64_t *result, void *stream);

/* ----------------------------------------------------------------- MiTCCRH --- */
/* Replaces MITCCRH.Hash (ot/mitccrh.go:93-128) over many keys at once: key
 * number g (renewKeys :70-89) is BE64(seed.D0 ^ g) || BE64(seed.D1); key
 * gid_start + i hashes blks[i*h .. i*h+h), in place: AES(x) ^ x. */
int gcb_mitccrh_hash(const gcb_label *seed, uint64_t gid_start, gcb_label *blks, uint64_t nkeys,
                     uint32_t h);
int gcb_mitccrh_hash_dev(const gcb_label *seed_host, uint64_t gid_start, gcb_label *blks,
                         uint64_t nkeys, uint32_t h, void *stream);

/* --------------------------------------------------- COT / ROT post-processing --- */
/* Replace the per-batch loops that follow the IKNP extension in COT.Send / COT.Receive
 * (ot/cot.go:157-181, :201-233) and ROT.Send / ROT.Receive (ot/rot.go:155-172, :192-197).
 * OT number j is hashed under MiTCCRH key j of NewMITCCRH(seed, otBatchSize) (ot/cot.go:47,
 * ot/mitccrh.go:61-128).  The _dev variants take device pointers (seed and delta stay host
 * pointers), so the IKNP output never leaves HBM; `stream` is a cudaStream_t.
 *   flags & GCB_COT_WIRE_BYTES: the message array is in the 16-byte SendLabel / ReceiveLabel
 *   encoding (BE64(D0) || BE64(D1), ot/label.go:105-114) instead of Go memory order, so it can
 *   be written to / read from the connection without a per-label conversion.
 * gcb_cot_send:    msgs[2j] = H_j(q[j]) ^ wires[j].L0, msgs[2j+1] = H_j(q[j] ^ delta) ^ wires[j].L1
 * gcb_cot_receive: result[j] = (choice[j] ? msgs[2j+1] : msgs[2j]) ^ H_j(t[j]); result may alias t
 * gcb_rot_send:    wires[j] = {H_j(q[j]), H_j(q[j] ^ delta)}
 * gcb_rot_receive: result[j] = H_j(t[j]); result may alias t */
#define GCB_COT_WIRE_BYTES 1u
int gcb_cot_send(const gcb_label *seed, const gcb_label *delta, const gcb_label *q, const gcb_wire *wires,
                 uint64_t n, gcb_label *msgs, uint32_t flags);
int gcb_cot_receive(const gcb_label *seed, const uint8_t *choice, const gcb_label *msgs, const gcb_label *t,
                    uint64_t n, gcb_label *result, uint32_t flags);
int gcb_rot_send(const gcb_label *seed, const gcb_label *delta, const gcb_label *q, uint64_t n, gcb_wire *wires);
int gcb_rot_receive(const gcb_label *seed, const gcb_label *t, uint64_t n, gcb_label *result);
int gcb_cot_send_dev(const gcb_label *seed, const gcb_label *delta, const gcb_label *q, const gcb_wire *wires,
                     uint64_t n, gcb_label *msgs, uint32_t flags, void *stream);
int gcb_cot_receive_dev(const gcb_label *seed, const uint8_t *choice, const gcb_label *msgs, const gcb_label *t,
                        uint64_t n, gcb_label *result, uint32_t flags, void *stream);
int gcb_rot_send_dev(const gcb_label *seed, const gcb_label *delta, const gcb_label *q, uint64_t n,
                     gcb_wire *wires, void *stream);
int gcb_rot_receive_dev(const gcb_label *seed, const gcb_label *t, uint64_t n, gcb_label *result, void *stream);

/* ------------------------------------------- IKNP malicious-mode consistency sums --- */
/* Replace the chi loops of IKNPSender.Send (ot/iknp.go:150-173) and IKNPReceiver.Receive
 * (:408-451), i.e. prgLabels + vectorInnPrdtSumNoRed / mul128 (ot/gf128.go:14-27,
 * ot/mul128_amd64.s): with chi_i = label number chi_start + i of the PRG keyed by seed2,
 *   out[0], out[1] = XOR_i mul128(chi_i, labels[i])   (low, high 128 bits, no reduction)
 *   out[2]         = XOR_{i : choice[i]} chi_i        (zero when choice is NULL: the sender)
 * The caller XORs the sums of its Send/Receive result (chi_start 0) and of the 256 extra
 * choice-vector OTs (chi_start n), then finishes the check as the reference does
 * (:175-191 / :453-465).  _dev: labels and choice are device pointers, out is a host pointer;
 * the call returns after the stream has produced the sums. */
int gcb_iknp_check_sums(const gcb_label *seed2, uint64_t chi_start, const gcb_label *labels,
                        const uint8_t *choice, uint64_t n, gcb_label out[3]);
int gcb_iknp_check_sums_dev(const gcb_label *seed2, uint64_t chi_start, const gcb_label *labels,
                            const uint8_t *choice, uint64_t n, gcb_label out[3], void *stream);

/* ------------------------------------------- garbled tables in the wire format --- */
/* The byte stream Garbler sends after Garble (circuit/garbler.go:69-82) and Evaluator parses
 * before Eval (circuit/evaluator.go:40-66): SendUint32(NumGates), then for every gate in file
 * order SendUint32(len(rows)) and the rows with SendLabel (16 bytes, BE64(D0) || BE64(D1)).
 * These calls convert between that stream and the dense row slab of gcb_garble / gcb_eval for
 * a whole batch, so the slab goes device -> pinned host -> socket with no per-label Go work
 * (and sha2pc/encoding.go:363-440 can copy the same bytes).
 *   wire bytes per instance = 4 + 4 * num_gates + 16 * num_rows   (gcb_tables_wire_size)
 *   stride: bytes between instances in the wire buffer; a multiple of 16 that is at least
 *   the wire size rounded up to 16, plus 16 for gcb_tables_from_wire (rows are read with
 *   aligned 32-byte windows).
 * gcb_tables_from_wire checks every count against the circuit ("wrong number of gates",
 * GCB_E_CORRUPT on a row count that does not match the gate type) on the host copy; the _dev
 * variants take device pointers, trust the counts and only move the rows. */
int gcb_tables_wire_size(const gcb_plan *plan, size_t *bytes);
int gcb_tables_to_wire(const gcb_plan *plan, uint32_t batch, const gcb_label *tables, uint8_t *dst, size_t stride);
int gcb_tables_from_wire(const gcb_plan *plan, uint32_t batch, const uint8_t *src, size_t stride, gcb_label *tables);
int gcb_tables_to_wire_dev(const gcb_plan *plan, uint32_t batch, const gcb_label *tables, uint8_t *dst,
                           size_t stride, void *stream);
int gcb_tables_from_wire_dev(const gcb_plan *plan, uint32_t batch, const uint8_t *src, size_t stride,
                             gcb_label *tables, void *stream);

/* --------------------------------------------------------- circuit front end --- */
/* For callers that are not Go: the reference's circuit file formats, statistics and plaintext
 * evaluation on the host (no device work), so that a file can become a plan and be checked
 * through this ABI alone.  Go callers keep circuit.Parse and pass Circuit.Gates to gcb_plan_create.
 *   gcb_circuit_parse     circuit.Parse on an in-memory file: GCB_FORMAT_BRISTOL follows
 *                         circuit/parser.go:265-494, GCB_FORMAT_MPCLC follows :71-211; the "wire
 *                         not set / not assigned" checks of the parser are applied to both
 *   gcb_circuit_from_gates  the same object from a gate array (levels and statistics computed)
 *   gcb_circuit_get_info  counts, IO.Size() of inputs / outputs, Stats[NumLevels] and
 *                         Stats[MaxWidth] of AssignLevels(TargetYao), circuit/circuit.go:206-254;
 *                         gcb_circuit_get_gates returns the gates with Level filled in
 *   gcb_circuit_compute   Circuit.Compute (circuit/computer.go:15-91) on wire bits:
 *                         in_bits [batch][num_inputs] of 0/1 -> out_bits [batch][num_outputs]
 *   gcb_circuit_plan      gcb_plan_create on the parsed gates */
typedef struct gcb_circuit gcb_circuit;
enum { GCB_FORMAT_BRISTOL = 0, GCB_FORMAT_MPCLC = 1 };
typedef struct {
    uint32_t num_gates, num_wires, num_inputs, num_outputs, num_input_args, num_output_args;
    uint32_t num_xor, num_xnor, num_and, num_or, num_inv;
    uint32_t num_levels, max_width;
} gcb_circuit_info;
int gcb_circuit_parse(const void *data, size_t len, int format, gcb_circuit **out);
int gcb_circuit_from_gates(const gcb_gate *gates, uint32_t num_gates, uint32_t num_wires, const uint32_t *inputs,
                           uint32_t n_input_args, const uint32_t *outputs, uint32_t n_output_args,
                           gcb_circuit **out);
void gcb_circuit_destroy(gcb_circuit *circ);
int gcb_circuit_get_info(const gcb_circuit *circ, gcb_circuit_info *info);
int gcb_circuit_get_gates(const gcb_circuit *circ, gcb_gate *gates);
int gcb_circuit_get_io(const gcb_circuit *circ, uint32_t *input_bits, uint32_t *output_bits);
int gcb_circuit_compute(const gcb_circuit *circ, uint32_t batch, const uint8_t *in_bits, uint8_t *out_bits);
int gcb_circuit_plan(const gcb_circuit *circ, gcb_plan **out);

#ifdef __cplusplus
}
#endif
#endif /* GCB200_H */
