/*
 * field_tmpl.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * 256-bit Montgomery prime field on 4 x u64 limbs, R = 2^256, instantiated
 * twice (Fr and Fq of BN254) by defining FNAME / FMOD before inclusion.
 * Follows halo2curves 0.3.1 `src/derive/field.rs` (field_arithmetic! macro:
 * `mul` = schoolbook 4x4 `mac` + `montgomery_reduce`, `add`/`sub` with a
 * conditional correction, `invert` = pow(p-2)) -- restated, the Rust sources
 * are not vendored in /root/reference (Cargo.lock:1329-1331).
 */
#define FCAT_(a, b) a##_##b
#define FCAT(a, b) FCAT_(a, b)
#define FN(name) FCAT(FNAME, name)

typedef struct { u64 l[4]; } FNAME;

static const u64 FN(MOD)[4] = FMOD;
static u64 FN(INV);          /* -p^-1 mod 2^64 */
static FNAME FN(R1);         /* R mod p  (= one)  */
static FNAME FN(R2);         /* R^2 mod p */
static FNAME FN(R3);         /* R^3 mod p */
static FNAME FN(ZERO_);

static inline int FN(geq_mod)(const u64 a[4]) {
  for (int i = 3; i >= 0; i--) {
    if (a[i] > FN(MOD)[i]) return 1;
    if (a[i] < FN(MOD)[i]) return 0;
  }
  return 1;
}
static inline void FN(sub_mod_raw)(u64 a[4]) {
  u128 b = 0;
  for (int i = 0; i < 4; i++) {
    u128 t = (u128)a[i] - FN(MOD)[i] - (u64)b;
    a[i] = (u64)t;
    b = (t >> 64) & 1;
  }
}
static inline FNAME FN(add)(FNAME a, FNAME b) {
  FNAME r; u128 c = 0;
  for (int i = 0; i < 4; i++) { c += (u128)a.l[i] + b.l[i]; r.l[i] = (u64)c; c >>= 64; }
  if (c || FN(geq_mod)(r.l)) FN(sub_mod_raw)(r.l);
  return r;
}
static inline FNAME FN(sub)(FNAME a, FNAME b) {
  FNAME r; u128 bw = 0;
  for (int i = 0; i < 4; i++) {
    u128 t = (u128)a.l[i] - b.l[i] - (u64)bw;
    r.l[i] = (u64)t; bw = (t >> 64) & 1;
  }
  if (bw) { u128 c = 0; for (int i = 0; i < 4; i++) { c += (u128)r.l[i] + FN(MOD)[i]; r.l[i] = (u64)c; c >>= 64; } }
  return r;
}
static inline FNAME FN(neg)(FNAME a) { return FN(sub)(FN(ZERO_), a); }
static inline FNAME FN(dbl)(FNAME a) { return FN(add)(a, a); }
static inline int FN(is_zero)(FNAME a) { return (a.l[0] | a.l[1] | a.l[2] | a.l[3]) == 0; }
static inline int FN(eq)(FNAME a, FNAME b) {
  return a.l[0] == b.l[0] && a.l[1] == b.l[1] && a.l[2] == b.l[2] && a.l[3] == b.l[3];
}

/* montgomery_reduce of an 8-limb value t < p * 2^256 */
static inline FNAME FN(mont_reduce)(u64 t[8]) {
  u64 carry2 = 0;
  for (int i = 0; i < 4; i++) {
    u64 k = t[i] * FN(INV);
    u128 c = 0;
    for (int j = 0; j < 4; j++) {
      c += (u128)k * FN(MOD)[j] + t[i + j];
      t[i + j] = (u64)c; c >>= 64;
    }
    u128 s = (u128)t[i + 4] + (u64)c + carry2;
    t[i + 4] = (u64)s; carry2 = (u64)(s >> 64);
  }
  FNAME r = {{t[4], t[5], t[6], t[7]}};
  if (carry2 || FN(geq_mod)(r.l)) FN(sub_mod_raw)(r.l);
  return r;
}
static inline FNAME FN(mul)(FNAME a, FNAME b) {
  u64 t[8] = {0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) {
      c += (u128)a.l[i] * b.l[j] + t[i + j];
      t[i + j] = (u64)c; c >>= 64;
    }
    t[i + 4] = (u64)c;
  }
  return FN(mont_reduce)(t);
}
static inline FNAME FN(sqr)(FNAME a) { return FN(mul)(a, a); }

/* canonical integer (LE limbs) -> Montgomery; value must be < p */
static inline FNAME FN(from_raw)(const u64 v[4]) {
  FNAME a = {{v[0], v[1], v[2], v[3]}};
  return FN(mul)(a, FN(R2));
}
/* Montgomery -> canonical LE limbs  (`to_repr`) */
static inline void FN(to_raw)(FNAME a, u64 out[4]) {
  u64 t[8] = {a.l[0], a.l[1], a.l[2], a.l[3], 0, 0, 0, 0};
  FNAME r = FN(mont_reduce)(t);
  memcpy(out, r.l, 32);
}
static inline FNAME FN(from_u64)(u64 v) { u64 t[4] = {v, 0, 0, 0}; return FN(from_raw)(t); }
/* from_u512 / from_bytes_wide: lo*R2 + hi*R3 (Montgomery) */
static inline FNAME FN(from_u512)(const u64 v[8]) {
  FNAME lo = {{v[0], v[1], v[2], v[3]}}, hi = {{v[4], v[5], v[6], v[7]}};
  return FN(add)(FN(mul)(lo, FN(R2)), FN(mul)(hi, FN(R3)));
}
static FNAME FN(pow)(FNAME base, const u64 e[4]) {
  FNAME acc = FN(R1);
  for (int i = 255; i >= 0; i--) {
    acc = FN(sqr)(acc);
    if ((e[i / 64] >> (i % 64)) & 1) acc = FN(mul)(acc, base);
  }
  return acc;
}
static FNAME FN(pow_u64)(FNAME base, u64 e) { u64 t[4] = {e, 0, 0, 0}; return FN(pow)(base, t); }
/* invert = a^(p-2); returns 0 for 0 */
static FNAME FN(inv)(FNAME a) {
  u64 e[4]; memcpy(e, FN(MOD), 32);
  e[0] -= 2; /* p is odd and its low limb is > 2 */
  return FN(pow)(a, e);
}
/* numeric order of canonical values (Fr: Ord) */
static inline int FN(cmp)(FNAME a, FNAME b) {
  u64 x[4], y[4]; FN(to_raw)(a, x); FN(to_raw)(b, y);
  for (int i = 3; i >= 0; i--) { if (x[i] != y[i]) return x[i] < y[i] ? -1 : 1; }
  return 0;
}
static void FN(batch_invert)(FNAME* v, size_t n) {
  if (!n) return;
  FNAME* pre = (FNAME*)malloc(n * sizeof(FNAME));
  FNAME acc = FN(R1);
  for (size_t i = 0; i < n; i++) { pre[i] = acc; if (!FN(is_zero)(v[i])) acc = FN(mul)(acc, v[i]); }
  acc = FN(inv)(acc);
  for (size_t i = n; i-- > 0;) {
    if (FN(is_zero)(v[i])) continue;
    FNAME t = FN(mul)(acc, pre[i]);
    acc = FN(mul)(acc, v[i]);
    v[i] = t;
  }
  free(pre);
}
static void FN(init)(void) {
  /* INV by Newton iteration on 2-adic inverse */
  u64 p0 = FN(MOD)[0], x = 1;
  for (int i = 0; i < 6; i++) x *= 2 - p0 * x;   /* x = p0^-1 mod 2^64 */
  FN(INV) = (u64)0 - x;
  /* R = 2^256 mod p by 256 modular doublings of 1 (plain, non-Montgomery add) */
  FNAME one = {{1, 0, 0, 0}};
  FNAME t = one;
  for (int i = 0; i < 256; i++) t = FN(add)(t, t);
  FN(R1) = t;
  for (int i = 0; i < 256; i++) t = FN(add)(t, t);
  FN(R2) = t;                 /* 2^512 mod p */
  FN(R3) = FN(mul)(FN(R2), FN(R2)); /* R^2 * R^2 / R = R^3 */
  memset(&FN(ZERO_), 0, sizeof(FNAME));
}

#undef FNAME
#undef FMOD
