// Host-side C API over csrc/field.cuh (host path) so that pytest can check the field/integer
// arithmetic against Python integers without a GPU.  The device path of the same header is checked
// on the GPU through the batch evaluator (tests/test_gpu_parity.py).
#include "../../circom-witnesscalc_b200/csrc/field.cuh"
#include <string.h>
using namespace gw;
static fe ld(const uint32_t* p) { fe r; memcpy(r.l, p, 32); return r; }
static void st(uint32_t* p, const fe& a) { memcpy(p, a.l, 32); }
extern "C" {
void t_mul(const uint32_t* a, const uint32_t* b, uint32_t* r) { st(r, fe_mul(ld(a), ld(b))); }
void t_mul_wide(const uint32_t* a, const uint32_t* b, uint32_t* r16) { u256_mul_wide(r16, a, b); }
void t_sqr_wide(const uint32_t* a, uint32_t* r16) { u256_sqr_wide(r16, a); }
void t_sqr(const uint32_t* a, uint32_t* r) { st(r, fe_sqr(ld(a))); }
void t_mul_lo(const uint32_t* a, const uint32_t* b, uint32_t* r8) { u256_mul_lo(r8, a, b); }
void t_mont_mul(const uint32_t* a, const uint32_t* b, uint32_t* r) { st(r, fe_mont_mul(ld(a), ld(b))); }
void t_to_mont(const uint32_t* a, uint32_t* r) { st(r, fe_to_mont(ld(a))); }
void t_from_mont(const uint32_t* a, uint32_t* r) { st(r, fe_from_mont(ld(a))); }
void t_add(const uint32_t* a, const uint32_t* b, uint32_t* r) { st(r, fe_add(ld(a), ld(b))); }
void t_sub(const uint32_t* a, const uint32_t* b, uint32_t* r) { st(r, fe_sub(ld(a), ld(b))); }
void t_neg(const uint32_t* a, uint32_t* r) { st(r, fe_neg(ld(a))); }
void t_reduce(const uint32_t* a, uint32_t* r) { st(r, fe_reduce256(ld(a))); }
void t_shr(const uint32_t* a, const uint32_t* b, uint32_t* r) { st(r, fe_shr(ld(a), ld(b))); }
int t_shl(const uint32_t* a, const uint32_t* b, uint32_t* r) { bool o; st(r, fe_shl(ld(a), ld(b), &o)); return o; }
int t_bitop(const uint32_t* a, const uint32_t* b, int which, uint32_t* r) { bool o; st(r, fe_bitop(ld(a), ld(b), which, &o)); return o; }
void t_bnot(const uint32_t* a, uint32_t* r) { st(r, fe_bnot(ld(a))); }
int t_cmp(const uint32_t* a, const uint32_t* b, int which) { return fe_cmp(ld(a), ld(b), which); }
void t_divrem(const uint32_t* a, const uint32_t* b, uint32_t* q, uint32_t* r) { fe qq, rr; u256_divrem(ld(a), ld(b), &qq, &rr); st(q, qq); st(r, rr); }
void t_pow(const uint32_t* a, const uint32_t* e, uint32_t* r) { st(r, fe_pow(ld(a), ld(e))); }
void t_inv(const uint32_t* a, uint32_t* r) { st(r, fe_inv_fermat(ld(a))); }
}
#include "../../circom-witnesscalc_b200/csrc/inv_safegcd.cuh"
extern "C" void t_inv_safegcd(const uint32_t* a, uint32_t* r) { st(r, gw::fe_inv(ld(a))); }
extern "C" void t_mul_hi_trunc(const uint32_t* a, const uint32_t* b, uint32_t* r16) { gw::u256_mul_hi_trunc(r16, a, b); }
// OP_DOT pieces: P (16 limbs) * 2^-256 mod M with n conditional subtractions, and one accumulated term
extern "C" void t_mont_reduce(const uint32_t* p16, int ncs, uint32_t* r) { gw::dot_acc A; gw::dot_load(A, p16); st(r, gw::fe_mont_reduce(A, ncs)); }
// n terms (kinds[t], xs[8t..], cs[8t..]) accumulated like the kernel does, then reduced
extern "C" void t_dot_eval(const uint32_t* kinds, const uint32_t* xs, const uint32_t* cs, int n, int ncs, uint32_t* r) {
  gw::dot_acc A; gw::dot_init(A);
  for (int t = 0; t < n; t++) gw::dot_term(A, kinds[t], ld(xs + 8 * t), ld(cs + 8 * t));
  st(r, gw::fe_mont_reduce(A, ncs));
}
extern "C" int t_emulates_ptx(void) {
#if defined(GW_EMULATE_PTX)
  return 1;
#else
  return 0;
#endif
}
