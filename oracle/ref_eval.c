/* CPU oracle, C restatement of the reference's witness evaluator.
 *
 * TEST INFRASTRUCTURE AND TIMED CPU BASELINE ONLY.  Nothing under circom-witnesscalc_b200/ links or
 * calls this file; it is used by tests/ (as the checker), by __graft_entry__.smoke() and by bench.py's
 * cpu_baseline / `--impl reference` legs.
 *
 * The reference (Rust, /root/reference) cannot be built in this image (no cargo/rustc, crates not
 * vendored), so this file restates its algorithm with the same cost structure:
 *   - values are ark-ff style Fr: 4 x 64-bit limbs in Montgomery form (ark-ff 0.4.2 / ark-bn254 0.4.0,
 *     Cargo.lock; not in the reference tree: restated from the published CIOS algorithm);
 *   - evaluate()            follows src/graph.rs:367-391 (one sequential loop over the nodes, then
 *                           into_bigint for every witness signal);
 *   - op semantics          follow src/graph.rs:102-144 (eval_fr), :188-197, :221-225, with the
 *                           Montgomery<->canonical hops exactly where :621-717 and :112-133 have them;
 *   - field inversion       is the binary extended Euclid that ark-ff uses for Fp::inverse;
 *   - the graph reader      follows src/storage.rs:214-249 and protos/messages.proto.
 * Favourable deviations from the reference (stated in every report): the graph is parsed once, not
 * per call (src/lib.rs:129-130), and nothing is printed.
 *
 * Parity pin: checked against the reference's unit-test vectors (src/graph.rs:779-883) and the
 * circuit-level known answers in tests/test_oracle_kat.py, and against oracle/pyoracle.py on random
 * graphs.  Where the reference panics (Shl overflow, Bor/Bxor == M, Pow, Id) this file follows the
 * circom semantics like the Python oracle's "circom" policy.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fr;   /* Montgomery form unless stated otherwise */

/* src/field.rs:3-4 and the constants ark-bn254 derives from it */
static const fr MOD = {{0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull}};
static const fr R1 = {{0xac96341c4ffffffbull, 0x36fc76959f60cd29ull, 0x666ea36f7879462eull, 0x0e0a77c19a07df2full}};   /* 2^256 mod M */
static const fr R2 = {{0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull}};   /* 2^512 mod M */
static const fr HALF = {{0xa1f0fac9f8000000ull, 0x9419f4243cdcb848ull, 0xdc2822db40c0ac2eull, 0x183227397098d014ull}}; /* src/graph.rs:720 */
static const uint64_t INV = 14042775128853446655ull;   /* -M^-1 mod 2^64, src/field.rs:6 */

static inline int big_is_zero(const fr* a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static inline int big_cmp(const fr* a, const fr* b) {
  for (int i = 3; i >= 0; i--) if (a->l[i] != b->l[i]) return a->l[i] < b->l[i] ? -1 : 1;
  return 0;
}
static inline uint64_t big_add(fr* r, const fr* a, const fr* b) {
  u128 c = 0;
  for (int i = 0; i < 4; i++) { c += (u128)a->l[i] + b->l[i]; r->l[i] = (uint64_t)c; c >>= 64; }
  return (uint64_t)c;
}
static inline uint64_t big_sub(fr* r, const fr* a, const fr* b) {
  uint64_t borrow = 0;
  for (int i = 0; i < 4; i++) {
    uint64_t t = a->l[i] - b->l[i], b1 = a->l[i] < b->l[i];
    uint64_t t2 = t - borrow, b2 = t < borrow;
    r->l[i] = t2; borrow = b1 | b2;
  }
  return borrow;
}
static inline void big_div2(fr* a) {
  a->l[0] = (a->l[0] >> 1) | (a->l[1] << 63); a->l[1] = (a->l[1] >> 1) | (a->l[2] << 63);
  a->l[2] = (a->l[2] >> 1) | (a->l[3] << 63); a->l[3] >>= 1;
}

/* Montgomery product (CIOS), result < M */
static inline fr fr_mul(const fr* a, const fr* b) {
  uint64_t t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) { c += (u128)a->l[j] * b->l[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
    c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
    uint64_t m = t[0] * INV;
    c = ((u128)m * MOD.l[0] + t[0]) >> 64;
    for (int j = 1; j < 4; j++) { c += (u128)m * MOD.l[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
    c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
  }
  fr r = {{t[0], t[1], t[2], t[3]}};
  fr s;
  if (!big_sub(&s, &r, &MOD)) r = s;
  return r;
}
static inline fr fr_add(const fr* a, const fr* b) { fr r, s; big_add(&r, a, b); if (!big_sub(&s, &r, &MOD)) r = s; return r; }
static inline fr fr_sub(const fr* a, const fr* b) { fr r; if (big_sub(&r, a, b)) big_add(&r, &r, &MOD); return r; }
/* Fr::new / from_bigint for a value already < M: canonical -> Montgomery */
static inline fr fr_from_canonical(const fr* c) { return fr_mul(c, &R2); }
/* into_bigint: Montgomery -> canonical */
static inline fr fr_into_bigint(const fr* a) { fr one = {{1, 0, 0, 0}}; return fr_mul(a, &one); }
/* any 256-bit integer -> Fr (Fr::new reduces mod M: graph.rs:376; from_le_bytes_mod_order: storage.rs:28) */
static fr fr_from_u256(fr v) {
  fr s;
  for (int k = 0; k < 5; k++) if (!big_sub(&s, &v, &MOD)) v = s;
  return fr_from_canonical(&v);
}
static inline fr fr_from_u64(uint64_t v) { fr c = {{v, 0, 0, 0}}; return fr_from_canonical(&c); }

/* Fp::inverse of ark-ff: binary extended Euclid on the Montgomery representation */
static fr fr_inverse(const fr* a) {
  fr one = {{1, 0, 0, 0}};
  fr u = *a, v = MOD, b = R2, c = {{0, 0, 0, 0}};
  while (big_cmp(&u, &one) != 0 && big_cmp(&v, &one) != 0) {
    while (!(u.l[0] & 1)) {
      big_div2(&u);
      if (b.l[0] & 1) { uint64_t cy = big_add(&b, &b, &MOD); big_div2(&b); b.l[3] |= cy << 63; } else big_div2(&b);
    }
    while (!(v.l[0] & 1)) {
      big_div2(&v);
      if (c.l[0] & 1) { uint64_t cy = big_add(&c, &c, &MOD); big_div2(&c); c.l[3] |= cy << 63; } else big_div2(&c);
    }
    if (big_cmp(&v, &u) < 0) { big_sub(&u, &u, &v); b = fr_sub(&b, &c); }
    else { big_sub(&v, &v, &u); c = fr_sub(&c, &b); }
  }
  return big_cmp(&u, &one) == 0 ? b : c;
}

/* 256-bit unsigned division (ruint's `/` and `%`, graph.rs:115,120): shift-subtract */
static void big_divrem(const fr* a, const fr* b, fr* q, fr* r) {
  fr quo = {{0, 0, 0, 0}}, rem = {{0, 0, 0, 0}}, t;
  for (int i = 255; i >= 0; i--) {
    rem.l[3] = (rem.l[3] << 1) | (rem.l[2] >> 63); rem.l[2] = (rem.l[2] << 1) | (rem.l[1] >> 63);
    rem.l[1] = (rem.l[1] << 1) | (rem.l[0] >> 63); rem.l[0] = (rem.l[0] << 1) | ((a->l[i >> 6] >> (i & 63)) & 1);
    if (!big_sub(&t, &rem, b)) { rem = t; quo.l[i >> 6] |= 1ull << (i & 63); }
  }
  *q = quo; *r = rem;
}
static fr big_shr(fr a, unsigned n) {      /* n < 256 */
  while (n >= 64) { a.l[0] = a.l[1]; a.l[1] = a.l[2]; a.l[2] = a.l[3]; a.l[3] = 0; n -= 64; }
  if (n) {
    a.l[0] = (a.l[0] >> n) | (a.l[1] << (64 - n)); a.l[1] = (a.l[1] >> n) | (a.l[2] << (64 - n));
    a.l[2] = (a.l[2] >> n) | (a.l[3] << (64 - n)); a.l[3] >>= n;
  }
  return a;
}
static fr big_shl(fr a, unsigned n) {      /* n < 256, truncated to 256 bits (BigInt::muln) */
  while (n >= 64) { a.l[3] = a.l[2]; a.l[2] = a.l[1]; a.l[1] = a.l[0]; a.l[0] = 0; n -= 64; }
  if (n) {
    a.l[3] = (a.l[3] << n) | (a.l[2] >> (64 - n)); a.l[2] = (a.l[2] << n) | (a.l[1] >> (64 - n));
    a.l[1] = (a.l[1] << n) | (a.l[0] >> (64 - n)); a.l[0] <<= n;
  }
  return a;
}

/* ---- node model (src/graph.rs:236-245) -------------------------------------------------------- */
enum { K_INPUT = 0, K_CONST = 1, K_UNO = 2, K_DUO = 3, K_TRES = 4 };
typedef struct { uint8_t kind, op; uint32_t a, b, c; } node_t;
typedef struct {
  node_t* nodes; uint64_t n_nodes;
  fr* consts; uint32_t n_consts;          /* Montgomery form (MontConstant) */
  uint32_t* witness; uint32_t n_witness;
  uint32_t n_inputs;                      /* get_inputs_size, src/lib.rs:138-152 */
  uint64_t n_ops;
} oracle_graph;

/* shift amount as graph.rs:621-672 reads it: 0 -> 0, >= 254 -> 254 (result 0), else the value */
static unsigned shift_amount(const fr* b_canon) {
  if (b_canon->l[1] | b_canon->l[2] | b_canon->l[3] || b_canon->l[0] >= 254) return 254;
  return (unsigned)b_canon->l[0];
}
static int is_neg(const fr* c) { return big_cmp(&HALF, c) < 0; }   /* graph.rs:724 */

static fr eval_duo(unsigned op, const fr* a, const fr* b) {
  static const fr zero = {{0, 0, 0, 0}};
  switch (op) {
    case 0: return fr_mul(a, b);                                              /* Mul  graph.rs:105 */
    case 1: { if (big_is_zero(b)) return zero; fr i = fr_inverse(b); return fr_mul(a, &i); }   /* Div :109 */
    case 2: return fr_add(a, b);                                              /* Add :110 */
    case 3: return fr_sub(a, b);                                              /* Sub :111 */
    case 4: {                                                                 /* Pow: unimplemented! in the reference (:141) */
      fr e = fr_into_bigint(b), r = R1;
      for (int i = 253; i >= 0; i--) { r = fr_mul(&r, &r); if ((e.l[i >> 6] >> (i & 63)) & 1) r = fr_mul(&r, a); }
      return r;
    }
    case 5: case 6: {                                                         /* Idiv :112-116, Mod :117-121 */
      if (big_is_zero(b)) return zero;
      fr ca = fr_into_bigint(a), cb = fr_into_bigint(b), q, r;
      big_divrem(&ca, &cb, &q, &r);
      return fr_from_canonical(op == 5 ? &q : &r);
    }
    case 7: case 8: {                                                         /* Eq :122-125, Neq :126-129 (Fr::cmp is canonical) */
      fr ca = fr_into_bigint(a), cb = fr_into_bigint(b);
      int eq = big_cmp(&ca, &cb) == 0;
      return (op == 7) == eq ? R1 : zero;
    }
    case 9: case 10: case 11: case 12: {                                      /* Lt Gt Leq Geq :130-133 -> :723-769 */
      fr ca = fr_into_bigint(a), cb = fr_into_bigint(b);
      int an = is_neg(&ca), bn = is_neg(&cb), c = big_cmp(&ca, &cb), r;
      if (an != bn) r = (op == 9 || op == 11) ? an : bn;
      else r = op == 9 ? c < 0 : op == 10 ? c > 0 : op == 11 ? c <= 0 : c >= 0;
      return fr_from_u64((uint64_t)r);
    }
    case 13: return (!big_is_zero(a) && !big_is_zero(b)) ? R1 : zero;         /* Land :134 */
    case 14: return (!big_is_zero(a) || !big_is_zero(b)) ? R1 : zero;         /* Lor :135 */
    case 15: {                                                                /* Shl :136 -> :621-635 */
      if (big_is_zero(b)) return *a;
      fr cb = fr_into_bigint(b);
      unsigned n = shift_amount(&cb);
      if (n >= 254) return zero;
      fr ca = fr_into_bigint(a);
      fr r = big_shl(ca, n);
      if (big_cmp(&r, &MOD) >= 0) { r.l[3] &= 0x3FFFFFFFFFFFFFFFull; fr s; if (!big_sub(&s, &r, &MOD)) r = s; }   /* reference panics here */
      return fr_from_canonical(&r);
    }
    case 16: {                                                                /* Shr :137 -> :637-672 */
      if (big_is_zero(b)) return *a;
      fr cb = fr_into_bigint(b);
      unsigned n = shift_amount(&cb);
      if (n >= 254) return zero;
      fr ca = fr_into_bigint(a);
      fr r = big_shr(ca, n);
      return fr_from_canonical(&r);
    }
    case 17: case 18: case 19: {                                              /* Bor Band Bxor :138-140 -> :674-717 */
      fr ca = fr_into_bigint(a), cb = fr_into_bigint(b), d, s;
      for (int i = 0; i < 4; i++) d.l[i] = op == 17 ? (ca.l[i] | cb.l[i]) : op == 18 ? (ca.l[i] & cb.l[i]) : (ca.l[i] ^ cb.l[i]);
      if (!big_sub(&s, &d, &MOD)) d = s;     /* d > M: d -= M; d == M (reference panics) -> 0 */
      return fr_from_canonical(&d);
    }
  }
  return zero;
}

static fr eval_uno(unsigned op, const fr* a) {
  static const fr zero = {{0, 0, 0, 0}};
  switch (op) {
    case 0: {                                                                 /* Neg graph.rs:190-194 */
      if (big_is_zero(a)) return zero;
      fr c = fr_into_bigint(a), x;
      big_sub(&x, &MOD, &c);
      return fr_from_canonical(&x);
    }
    case 1: return *a;                                                        /* Id: unimplemented! (:195) */
    case 2: return big_is_zero(a) ? R1 : zero;                                /* Lnot (extension) */
    case 3: {                                                                 /* Bnot (extension) */
      fr c = fr_into_bigint(a), s;
      for (int i = 0; i < 4; i++) c.l[i] = ~c.l[i];
      c.l[3] &= 0x3FFFFFFFFFFFFFFFull;
      if (!big_sub(&s, &c, &MOD)) c = s;
      return fr_from_canonical(&c);
    }
  }
  return zero;
}

/* graph::evaluate, src/graph.rs:367-391.  inputs: I x 32 B LE; out: W x 32 B LE canonical; values: scratch of n_nodes */
static void evaluate_one(const oracle_graph* g, const uint8_t* inputs, uint8_t* out, fr* values) {
  const node_t* nd = g->nodes;
  for (uint64_t i = 0; i < g->n_nodes; i++, nd++) {
    switch (nd->kind) {
      case K_CONST: values[i] = g->consts[nd->a]; break;                                   /* :375 */
      case K_INPUT: { fr v; memcpy(&v, inputs + 32 * (size_t)nd->a, 32); values[i] = fr_from_u256(v); break; }   /* :376 */
      case K_DUO: values[i] = eval_duo(nd->op, &values[nd->a], &values[nd->b]); break;    /* :377 */
      case K_UNO: values[i] = eval_uno(nd->op, &values[nd->a]); break;                    /* :378 */
      default: values[i] = big_is_zero(&values[nd->a]) ? values[nd->c] : values[nd->b]; break;   /* TernCond :379, :221-225 */
    }
  }
  for (uint32_t j = 0; j < g->n_witness; j++) {                                            /* :385-388 */
    fr c = fr_into_bigint(&values[g->witness[j]]);
    memcpy(out + 32 * (size_t)j, &c, 32);
  }
}

/* ---- graph file reader (src/storage.rs:214-249, protos/messages.proto) -------------------------- */
typedef struct { const uint8_t* p; const uint8_t* end; int err; } rd_t;
static uint64_t rd_varint(rd_t* r) {
  uint64_t v = 0; int sh = 0;
  for (;;) {
    if (r->p >= r->end || sh > 63) { r->err = 1; return 0; }
    uint8_t b = *r->p++;
    v |= (uint64_t)(b & 0x7F) << sh;
    if (!(b & 0x80)) return v;
    sh += 7;
  }
}
static rd_t rd_sub(rd_t* r) {
  uint64_t n = rd_varint(r);
  rd_t s = {r->p, r->p, r->err};
  if (r->err || n > (uint64_t)(r->end - r->p)) { r->err = 1; s.err = 1; return s; }
  s.end = r->p + n; r->p += n;
  return s;
}
static void rd_skip(rd_t* r, unsigned wt) {
  if (wt == 0) rd_varint(r);
  else if (wt == 2) rd_sub(r);
  else if (wt == 1 && r->end - r->p >= 8) r->p += 8;
  else if (wt == 5 && r->end - r->p >= 4) r->p += 4;
  else r->err = 1;
}
static void rd_uints(rd_t m, uint32_t* f, unsigned nmax, int* err) {
  for (unsigned i = 0; i <= nmax; i++) f[i] = 0;
  while (m.p < m.end && !m.err) {
    uint64_t key = rd_varint(&m);
    if ((key & 7) == 0 && (key >> 3) >= 1 && (key >> 3) <= nmax) f[key >> 3] = (uint32_t)rd_varint(&m);
    else rd_skip(&m, (unsigned)(key & 7));
  }
  if (m.err) *err = 1;
}

void oracle_free(oracle_graph* g) {
  if (!g) return;
  free(g->nodes); free(g->consts); free(g->witness); free(g);
}

oracle_graph* oracle_load(const uint8_t* data, size_t len) {
  static const char magic[] = "wtns.graph.001";                 /* storage.rs:16 */
  if (len < 22 || memcmp(data, magic, 14) != 0) return NULL;
  uint64_t n = 0;
  for (int i = 0; i < 8; i++) n |= (uint64_t)data[14 + i] << (8 * i);   /* u64 LE, storage.rs:228 */
  if (n > len) return NULL;
  oracle_graph* g = (oracle_graph*)calloc(1, sizeof *g);
  g->nodes = (node_t*)calloc(n ? n : 1, sizeof(node_t));
  g->consts = (fr*)calloc(n ? n : 1, sizeof(fr));
  g->n_nodes = n;
  rd_t f = {data + 22, data + len, 0};
  for (uint64_t i = 0; i < n && !f.err; i++) {
    rd_t msg = rd_sub(&f);
    node_t nd; memset(&nd, 0, sizeof nd);
    int have = 0;
    while (msg.p < msg.end && !msg.err) {
      uint64_t key = rd_varint(&msg);
      unsigned fno = (unsigned)(key >> 3), wt = (unsigned)(key & 7);
      if (wt != 2 || fno < 1 || fno > 5) { rd_skip(&msg, wt); continue; }
      rd_t in = rd_sub(&msg);
      uint32_t u[5];
      have = 1;
      memset(&nd, 0, sizeof nd);
      if (fno == 1) { rd_uints(in, u, 1, &f.err); nd.kind = K_INPUT; nd.a = u[1]; }
      else if (fno == 2) {
        fr v = {{0, 0, 0, 0}}; int got = 0;
        while (in.p < in.end && !in.err) {
          uint64_t k2 = rd_varint(&in);
          if ((k2 >> 3) == 1 && (k2 & 7) == 2) {
            rd_t big = rd_sub(&in);
            got = 1;
            while (big.p < big.end && !big.err) {
              uint64_t k3 = rd_varint(&big);
              if ((k3 >> 3) == 1 && (k3 & 7) == 2) {
                rd_t by = rd_sub(&big);
                size_t bl = (size_t)(by.end - by.p);
                if (bl > 32) { f.err = 1; break; }            /* the serializer never writes more than 32 bytes */
                uint8_t buf[32] = {0};
                memcpy(buf, by.p, bl);
                memcpy(&v, buf, 32);
              } else rd_skip(&big, (unsigned)(k3 & 7));
            }
            if (big.err) f.err = 1;
          } else rd_skip(&in, (unsigned)(k2 & 7));
        }
        if (!got || in.err) f.err = 1;
        nd.kind = K_CONST; nd.a = g->n_consts;
        g->consts[g->n_consts++] = fr_from_u256(v);            /* MontConstant, storage.rs:26-29 */
      }
      else if (fno == 3) { rd_uints(in, u, 2, &f.err); nd.kind = K_UNO; nd.op = (uint8_t)u[1]; nd.a = u[2]; if (u[1] > 3) f.err = 1; }
      else if (fno == 4) { rd_uints(in, u, 3, &f.err); nd.kind = K_DUO; nd.op = (uint8_t)u[1]; nd.a = u[2]; nd.b = u[3]; if (u[1] > 19) f.err = 1; }
      else { rd_uints(in, u, 4, &f.err); nd.kind = K_TRES; nd.op = (uint8_t)u[1]; nd.a = u[2]; nd.b = u[3]; nd.c = u[4]; if (u[1] > 0) f.err = 1; }
    }
    if (!have || msg.err) f.err = 1;
    if (nd.kind >= K_UNO && nd.a >= i) f.err = 1;               /* graph.rs:343-356 */
    if (nd.kind >= K_DUO && nd.b >= i) f.err = 1;
    if (nd.kind == K_TRES && nd.c >= i) f.err = 1;
    if (nd.kind >= K_UNO) g->n_ops++;
    g->nodes[i] = nd;
  }
  rd_t meta = rd_sub(&f);
  uint32_t cap = 16; g->witness = (uint32_t*)malloc(cap * 4);
  uint32_t max_slot = 0;
  while (meta.p < meta.end && !meta.err) {
    uint64_t key = rd_varint(&meta);
    unsigned fno = (unsigned)(key >> 3), wt = (unsigned)(key & 7);
    if (fno == 1 && (wt == 2 || wt == 0)) {
      rd_t pk = wt == 2 ? rd_sub(&meta) : meta;
      do {
        uint32_t s = (uint32_t)rd_varint(wt == 2 ? &pk : &meta);
        if (g->n_witness == cap) { cap *= 2; g->witness = (uint32_t*)realloc(g->witness, cap * 4); }
        if (s >= n) f.err = 1;
        g->witness[g->n_witness++] = s;
      } while (wt == 2 && pk.p < pk.end && !pk.err);
      if (wt == 2 && pk.err) f.err = 1;
    } else if (fno == 2 && wt == 2) {
      rd_t e = rd_sub(&meta);
      while (e.p < e.end && !e.err) {
        uint64_t k2 = rd_varint(&e);
        if ((k2 >> 3) == 2 && (k2 & 7) == 2) { uint32_t u[3]; rd_uints(rd_sub(&e), u, 2, &f.err); if (u[2] && u[1] + u[2] - 1 > max_slot) max_slot = u[1] + u[2] - 1; }
        else rd_skip(&e, (unsigned)(k2 & 7));
      }
    } else rd_skip(&meta, wt);
  }
  if (meta.err) f.err = 1;
  for (uint64_t i = 0; i < n; i++) if (g->nodes[i].kind == K_INPUT && g->nodes[i].a > max_slot) max_slot = g->nodes[i].a;
  g->n_inputs = max_slot + 1;
  if (f.err) { oracle_free(g); return NULL; }
  return g;
}

void oracle_info(const oracle_graph* g, uint64_t* out4) {
  out4[0] = g->n_nodes; out4[1] = g->n_inputs; out4[2] = g->n_witness; out4[3] = g->n_ops;
}

typedef struct { const oracle_graph* g; const uint8_t* in; uint8_t* out; size_t lo, hi; } job_t;
static void* worker(void* arg) {
  job_t* j = (job_t*)arg;
  const oracle_graph* g = j->g;
  fr* values = (fr*)malloc((g->n_nodes ? g->n_nodes : 1) * sizeof(fr));
  for (size_t w = j->lo; w < j->hi; w++) {
    uint8_t row[32];
    /* slot 0 of the input buffer is forced to 1 (get_inputs_buffer, src/lib.rs:177-181) */
    const uint8_t* in = j->in + w * (size_t)g->n_inputs * 32;
    uint8_t* tmp = NULL;
    memset(row, 0, 32); row[0] = 1;
    if (memcmp(in, row, 32) != 0) { tmp = (uint8_t*)malloc((size_t)g->n_inputs * 32); memcpy(tmp, in, (size_t)g->n_inputs * 32); memcpy(tmp, row, 32); in = tmp; }
    evaluate_one(g, in, j->out + w * (size_t)g->n_witness * 32, values);
    free(tmp);
  }
  free(values);
  return NULL;
}

/* inputs: B x I x 32 B LE, out: B x W x 32 B LE; one witness per thread at a time, static partition */
int oracle_evaluate_batch(const oracle_graph* g, const uint8_t* inputs, size_t B, uint8_t* out, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  if ((size_t)n_threads > B) n_threads = B ? (int)B : 1;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * n_threads);
  job_t* jobs = (job_t*)malloc(sizeof(job_t) * n_threads);
  for (int t = 0; t < n_threads; t++) {
    jobs[t].g = g; jobs[t].in = inputs; jobs[t].out = out;
    jobs[t].lo = B * t / n_threads; jobs[t].hi = B * (t + 1) / n_threads;
    if (t > 0) pthread_create(&th[t], NULL, worker, &jobs[t]);
  }
  worker(&jobs[0]);
  for (int t = 1; t < n_threads; t++) pthread_join(th[t], NULL);
  free(th); free(jobs);
  return 0;
}

/* unit-test hooks: one op on canonical operands (Montgomery conversion inside, like Fr::from / into_bigint) */
void oracle_op_duo(unsigned op, const uint8_t* a32, const uint8_t* b32, uint8_t* r32) {
  fr a, b; memcpy(&a, a32, 32); memcpy(&b, b32, 32);
  a = fr_from_u256(a); b = fr_from_u256(b);
  fr r = eval_duo(op, &a, &b);
  r = fr_into_bigint(&r);
  memcpy(r32, &r, 32);
}
void oracle_op_uno(unsigned op, const uint8_t* a32, uint8_t* r32) {
  fr a; memcpy(&a, a32, 32);
  a = fr_from_u256(a);
  fr r = eval_uno(op, &a);
  r = fr_into_bigint(&r);
  memcpy(r32, &r, 32);
}
