"""Builds tests/golden/ref_proof_points.json from the reference's own golden proof (run in the build container only).

bin/assets/proof.bin (consumed by the reference's test_cli_verify, bin/src/main.rs:270-279) is a halo2-axiom / snark-verifier
proof: 47 chunks of 32 bytes, commitments in halo2curves >= 0.4 compressed form.  Without ./params/{vk.bin, *.srs} (not
shipped) the proof cannot be verified, but its bytes still pin the POINT ENCODING of that stack: six chunks carry bit 6 of
byte 31, which is only a valid G1 encoding if that bit is a flag outside the 254-bit x coordinate (with the flag left in,
x >= p).  The y coordinates below are derived with Python integers under the rule "bit 6 = parity of y".

    python tests/golden/make_ref_proof_points.py
"""
import json
import os

P = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
R = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
SRC = "/root/reference/bin/assets/proof.bin"


def main():
    text = open(SRC).read().strip()
    raw = bytes.fromhex(text[2:] if text.startswith("0x") else text)
    assert len(raw) % 32 == 0
    chunks = [raw[i:i + 32] for i in range(0, len(raw), 32)]
    points, flagged = {}, []
    for i, c in enumerate(chunks):
        v = int.from_bytes(c, "little")
        assert v >> 255 == 0, "identity flag never set in this proof"
        sign, x = (v >> 254) & 1, v & ((1 << 254) - 1)
        if sign:
            flagged.append(i)
        if x >= P:
            continue
        y = pow((x ** 3 + 3) % P, (P + 1) // 4, P)
        if (y * y - x ** 3 - 3) % P:
            assert not sign, "a flagged chunk must be a point"
            continue
        if (y & 1) != sign:
            y = P - y
        # unflagged chunks below r may equally be scalars (evaluations); only the flagged ones are certainly points
        points[str(i)] = ["%064x" % x, "%064x" % y]
    assert flagged and all(str(i) in points for i in flagged)
    out = {"source": "bin/assets/proof.bin of the reference (bin/src/main.rs:270-279)",
           "encoding": "halo2curves >= 0.4 compressed: x little-endian, parity of y in bit 6 of byte 31, identity flag in bit 7",
           "chunks": [c.hex() for c in chunks], "flagged": flagged, "points": points}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_proof_points.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("%d chunks, %d flagged, %d decodable" % (len(chunks), len(flagged), len(points)))


if __name__ == "__main__":
    main()
