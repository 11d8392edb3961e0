"""vk.bin / pk.bin: the byte formats of halo2's `VerifyingKey::{write, read}` and `ProvingKey::{write, read}` in the two raw
SerdeFormats (the reference writes and reads both files with `SerdeFormat::RawBytesUnchecked`: crates/p256-ecdsa/src/base.rs:145-153
`vk.write(&mut file, SerdeFormat::RawBytesUnchecked)`, gen_pk's `params/pk.bin`; bin/src/main.rs:205-225 `ProvingKey::from_bytes`,
`VerifyingKey::from_bytes`).

[UPSTREAM?] The format lives in the un-vendored halo2 crates (halo2_proofs plonk.rs / permutation/keygen.rs / poly.rs, helpers.rs;
halo2-axiom for the p256-ecdsa stack) and is restated here from memory; the reference ships neither file (`./params` is
git-ignored), so nothing in its tree can pin these bytes -- PARITY UNPINNED.  What the tests show is self-consistency (write -> read
round trips, a key read back from bytes verifies / evaluates like the original).

Layout, all counts as big-endian u32:
  vk :=  k | #fixed_commitments | fixed_commitments[] | permutation_commitments[#permutation columns] | selectors[#selectors]
         (each selector: n bits packed 8 per byte, least significant bit first)
  pk :=  vk | l0 | l_last | l_active_row | fixed_values[] | fixed_polys[] | fixed_cosets[] | permutations[] | polys[] | cosets[]
         polynomial := #values | values[];  polynomial slice := #polynomials | polynomial[]
Raw formats: a field element is its four 64-bit Montgomery limbs, little-endian (= this package's array layout, so reading is a view,
not a conversion); a G1 point is x then y in that form, the identity (0, 0).  `raw_bytes` additionally checks that every limb vector
is below the modulus and every point on the curve; `raw_bytes_unchecked` does not.  The `processed` format (compressed points,
canonical scalars) is the SRS file's business (ParamsKZG.read); key files in it are not supported here.

`transcript_repr` is NOT in the file: halo2 recomputes it from the Debug rendering of the pinned key, which cannot be restated; pass
the one your circuit description uses.
"""
import struct

import numpy as np

R_MOD = 0x30644e72e131a029b85045b68181585d2833e84879b97091_43e1f593f0000001
P_MOD = 0x30644e72e131a029b85045b68181585d97816a916871ca8d_3c208c16d87cfd47
FORMATS = ("raw_bytes", "raw_bytes_unchecked")


class KeyFileError(ValueError):
    pass


def _fmt(fmt):
    if fmt not in FORMATS:
        raise KeyFileError("key files are read / written in %s (got %r)" % (" or ".join(FORMATS), fmt))
    return fmt == "raw_bytes"


def _limbs_below(arr, modulus):
    """every row of the (m, 4) u64 array < modulus"""
    m = [(modulus >> (64 * i)) & ((1 << 64) - 1) for i in range(4)]
    lt = np.zeros(arr.shape[0], dtype=bool)
    eq = np.ones(arr.shape[0], dtype=bool)
    for i in (3, 2, 1, 0):
        lt |= eq & (arr[:, i] < np.uint64(m[i]))
        eq &= arr[:, i] == np.uint64(m[i])
    return bool(lt.all())


def _check_points(pts):
    """(m, 8) Montgomery affine points: coordinates canonical and y^2 = x^3 + 3 (or the identity)"""
    if not _limbs_below(pts[:, :4], P_MOD) or not _limbs_below(pts[:, 4:], P_MOD):
        raise KeyFileError("point coordinate not below the field modulus")
    rinv = pow(1 << 256, -1, P_MOD)
    for row in pts:
        x = sum(int(row[i]) << (64 * i) for i in range(4)) * rinv % P_MOD
        y = sum(int(row[4 + i]) << (64 * i) for i in range(4)) * rinv % P_MOD
        if (x or y) and (y * y - x * x * x - 3) % P_MOD:
            raise KeyFileError("point not on the curve")


class _Reader:
    def __init__(self, data, checked):
        self.buf = memoryview(bytes(data)) if not isinstance(data, memoryview) else data
        self.pos, self.checked = 0, checked

    def take(self, n):
        if self.pos + n > len(self.buf):
            raise KeyFileError("key file too short")
        out = self.buf[self.pos:self.pos + n]
        self.pos += n
        return out

    def u32(self):
        return struct.unpack(">I", self.take(4))[0]

    def scalars(self, count):
        arr = np.frombuffer(self.take(32 * count), dtype="<u8").reshape(count, 4)
        if self.checked and count and not _limbs_below(arr, R_MOD):
            raise KeyFileError("scalar not below the field modulus")
        return arr

    def points(self, count):
        arr = np.frombuffer(self.take(64 * count), dtype="<u8").reshape(count, 8)
        if self.checked and count:
            _check_points(arr)
        return arr

    def polynomial(self, expect=None):
        n = self.u32()
        if expect is not None and n != expect:
            raise KeyFileError("polynomial of %d values where %d were expected" % (n, expect))
        return self.scalars(n)

    def polynomial_slice(self, expect=None):
        return [self.polynomial(expect) for _ in range(self.u32())]


def _polynomial(arr):
    a = np.ascontiguousarray(arr, dtype="<u8").reshape(-1, 4)
    return struct.pack(">I", a.shape[0]) + a.tobytes()


def _polynomial_slice(polys):
    return struct.pack(">I", len(polys)) + b"".join(_polynomial(p) for p in polys)


def write_vk(vk, fmt="raw_bytes_unchecked"):
    """vk: {"k", "fixed_commitments": (F, 8) u64, "permutation_commitments": (P, 8) u64, "selectors": [bool (n,)] (optional)}"""
    _fmt(fmt)
    fc = np.ascontiguousarray(vk["fixed_commitments"], dtype="<u8").reshape(-1, 8)
    pc = np.ascontiguousarray(vk["permutation_commitments"], dtype="<u8").reshape(-1, 8)
    out = [struct.pack(">I", int(vk["k"])), struct.pack(">I", fc.shape[0]), fc.tobytes(), pc.tobytes()]
    for sel in vk.get("selectors", []):
        sel = np.asarray(sel, dtype=bool)
        if sel.shape != (1 << int(vk["k"]),):
            raise KeyFileError("selector of %s rows in a key of 2^%d" % (sel.shape, vk["k"]))
        out.append(np.packbits(sel, bitorder="little").tobytes())
    return b"".join(out)


def _read_vk(rd, num_permutation_columns, num_selectors):
    k = rd.u32()
    if k > 30:
        raise KeyFileError("k = %d" % k)
    nf = rd.u32()
    vk = {"k": k, "fixed_commitments": rd.points(nf), "permutation_commitments": rd.points(num_permutation_columns), "selectors": []}
    n = 1 << k
    for _ in range(num_selectors):
        bits = np.unpackbits(np.frombuffer(rd.take((n + 7) // 8), dtype=np.uint8), bitorder="little")[:n]
        vk["selectors"].append(bits.astype(bool))
    return vk


def read_vk(data, num_permutation_columns, num_selectors=0, fmt="raw_bytes_unchecked"):
    """VerifyingKey::read: the column counts come from the circuit (`cs.permutation.columns.len()`, `cs.num_selectors`), as upstream
    takes them from `ConcreteCircuit::configure`.  Raises KeyFileError on trailing bytes."""
    rd = _Reader(data, _fmt(fmt))
    vk = _read_vk(rd, num_permutation_columns, num_selectors)
    if rd.pos != len(rd.buf):
        raise KeyFileError("%d trailing bytes after the verifying key" % (len(rd.buf) - rd.pos))
    return vk


PK_SINGLE = ("l0", "l_last", "l_active_row")
PK_SLICES = ("fixed_values", "fixed_polys", "fixed_cosets", "permutations", "permutation_polys", "permutation_cosets")


def write_pk(pk, fmt="raw_bytes_unchecked"):
    """pk: {"vk": (see write_vk), "l0", "l_last", "l_active_row": (ext_n, 4) u64, and the six polynomial lists of PK_SLICES}"""
    out = [write_vk(pk["vk"], fmt)]
    out += [_polynomial(pk[name]) for name in PK_SINGLE]
    out += [_polynomial_slice(pk[name]) for name in PK_SLICES]
    return b"".join(out)


def read_pk(data, num_permutation_columns, num_selectors=0, fmt="raw_bytes_unchecked", extended_k=None):
    """ProvingKey::read.  With `extended_k` given, polynomial lengths are checked against n = 2^k and 2^extended_k."""
    rd = _Reader(data, _fmt(fmt))
    vk = _read_vk(rd, num_permutation_columns, num_selectors)
    n = 1 << vk["k"]
    en = (1 << extended_k) if extended_k is not None else None
    pk = {"vk": vk}
    for name in PK_SINGLE:
        pk[name] = rd.polynomial(en)
    want = {"fixed_values": n, "fixed_polys": n, "fixed_cosets": en, "permutations": n, "permutation_polys": n, "permutation_cosets": en}
    for name in PK_SLICES:
        pk[name] = rd.polynomial_slice(want[name] if extended_k is not None else None)
    if len(pk["fixed_values"]) != vk["fixed_commitments"].shape[0] or len(pk["fixed_polys"]) != len(pk["fixed_values"]) \
            or len(pk["fixed_cosets"]) != len(pk["fixed_values"]):
        raise KeyFileError("fixed column counts disagree inside the key")
    if any(len(pk[s]) != num_permutation_columns for s in ("permutations", "permutation_polys", "permutation_cosets")):
        raise KeyFileError("permutation column counts disagree with the circuit")
    if rd.pos != len(rd.buf):
        raise KeyFileError("%d trailing bytes after the proving key" % (len(rd.buf) - rd.pos))
    return pk


PINNING_KEYS = ("k", "num_advice_per_phase", "num_fixed", "num_lookup_advice_per_phase", "lookup_bits", "num_instance_columns")


def read_pinning(text):
    """./params/pinning.json (bin/src/main.rs:195-198; written at crates/p256-ecdsa/src/base.rs:155-164 as
    `serde_json::to_writer_pretty(&(circuit.params(), circuit.break_points()))`): a two-element JSON array -- halo2-base's
    BaseCircuitParams [UPSTREAM?: field names from memory] and the per-phase thread break points.  Returns (params dict, break
    points); the circuit those parameters configure is halo2-base's, which this backend does not rebuild."""
    import json
    doc = json.loads(text)
    if not isinstance(doc, list) or len(doc) != 2 or not isinstance(doc[0], dict) or not isinstance(doc[1], list):
        raise KeyFileError("pinning.json: expected [BaseCircuitParams, break points]")
    params, breaks = doc
    missing = [key for key in PINNING_KEYS if key not in params]
    if missing:
        raise KeyFileError("pinning.json: BaseCircuitParams without %s" % ", ".join(missing))
    if not all(isinstance(phase, list) and all(isinstance(v, int) for v in phase) for phase in breaks):
        raise KeyFileError("pinning.json: break points are lists of row indices per phase")
    return params, breaks
