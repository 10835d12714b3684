#!/usr/bin/env python3
"""Generate tests/golden/*.json.  Run in the build container, where /root/reference exists:

    python tests/golden/make_golden.py

1. reference_poseidon_kats.json — values COPIED OUT OF THE REFERENCE'S OWN TEST FIXTURES (hex strings and
   u64 literals, no code), each with the file:line it came from.  These are the only results the reference
   pins on the hot path: all of them are Poseidon-Goldilocks outputs.
2. commit_vectors.json — end-to-end commitment vectors produced by the oracle (oracle/oracle.c) on
   SplitMix64 inputs.  The reference holds no NTT / LDE / cap fixture, so these are oracle-derived
   ("parity unpinned"); they freeze the oracle's behaviour and are cross-checked against SURVEY.md App. C,
   which was derived independently from O(n^2) definitions.
3. opening_proof_vectors.json — opening proofs (prove_openings / fri_proof) produced by the CPU restatement
   oracle/fri_ref.py on SplitMix64 coefficients, accepted by its verifier at generation time.  Oracle-derived as well
   ("parity unpinned": the reference holds no proof fixture); they freeze the restatement and are re-checked on the GPU.
Nothing under tests/ reads /root/reference at run time; only this generator does.
"""
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def u64_literals(path, lo, hi):
    with open(os.path.join(REF, path)) as f:
        lines = f.readlines()[lo - 1:hi]
    return [int(x) for x in re.findall(r"from_canonical_u64\((\d+)\)", "".join(lines))]


def reference_kats():
    out = {"_comment": "values copied from the reference's fixtures; see 'source' on each entry"}
    v = u64_literals("src/transaction/circuits/mod.rs", 211, 218)
    assert len(v) == 4
    out["two_to_one_zero_zero"] = {"source": "src/transaction/circuits/mod.rs:211-218 (default tx_hash)", "elements": v}

    with open(os.path.join(REF, "src/rollup/circuits/mod.rs")) as f:
        line = f.readlines()[103]
    m = re.search(r'let encoded_block_detail = "(.*)";', line)
    blob = json.loads(m.group(1).encode().decode("unicode_escape"))
    sib = blob["block_headers_proof_siblings"]
    assert len(sib) == 32
    out["zero_hash_chain"] = {"source": "src/rollup/circuits/mod.rs:104 block_headers_proof_siblings (s0=0, s_{i+1}=H(s_i,s_i))",
                              "hex": sib}
    hdr = blob["prev_block_header"]
    out["prev_block_header"] = {"source": "src/rollup/circuits/mod.rs:104 prev_block_header (BlockHeader::new(4))",
                                "block_headers_digest": hdr["block_headers_digest"],
                                "transactions_digest": hdr["transactions_digest"],
                                "deposit_digest": hdr["deposit_digest"]}

    sk1 = u64_literals("src/bin/block_circuit.rs", 81, 88)
    sk2 = u64_literals("src/bin/block_circuit.rs", 157, 164)
    n1 = u64_literals("src/bin/block_circuit.rs", 284, 291)
    n2 = u64_literals("src/bin/block_circuit.rs", 316, 323)
    with open(os.path.join(REF, "test_cases/block1_info.json")) as f:
        info = json.load(f)
    out["accounts"] = {
        "source": "src/bin/block_circuit.rs:81-88,157-164 (private keys) + test_cases/block1_info.json address_list; "
                  "address = two_to_one(sk, sk) (src/zkdsa/account.rs:164-170)",
        "sk": [sk1, sk2],
        "address_hex": [a["sender_address"] for a in info["address_list"]],
    }
    out["tx_hashes"] = {
        "source": "test_cases/block1_info.json transactions; tx = two_to_one(diff_root, nonce), nonces at "
                  "src/bin/block_circuit.rs:284-291,316-323; diff_root = layered SMT of keys at :108-123 "
                  "(leaf hash = hash_pad([key,value,1]), src/sparse_merkle_tree/goldilocks_poseidon/mod.rs:161-183)",
        "nonces": [n1, n2],
        "smt_entries": [{"key": [407, 305, 8012], "value": 2053}, {"key": [832, 471, 8012], "value": 1111}],
        "tx_hex": info["transactions"],
    }
    return out


def commit_vectors():
    import numpy as np
    from oracle import oracle as O
    cases = []
    for (n_log, k, r, h, coeffs, salted) in [(3, 9, 3, 2, False, False), (2, 3, 1, 0, False, False),
                                             (4, 5, 3, 4, False, False), (5, 135, 3, 4, False, False),
                                             (4, 16, 3, 4, True, False), (3, 20, 3, 4, False, True),
                                             (0, 7, 3, 3, False, False), (6, 2, 1, 7, False, False)]:
        n = 1 << n_log
        v = O.synthetic_values(k, n)
        salt = None
        if salted:
            salt = O.synthetic_values(4, n << r, seed=7)
        res = O.commit(v, r, h, is_coeffs=coeffs, salt=salt)
        N = n << r
        cases.append({
            "n_log": n_log, "k": k, "rate_bits": r, "cap_height": h, "is_coeffs": coeffs, "salt_seed": 7 if salted else None,
            "input": "v[c][i] = splitmix64(c*n + i) mod p (SURVEY.md App. C)",
            "cap": [[int(x) for x in row] for row in res["cap"]],
            "coeffs_col0_head": [int(x) for x in res["coeffs"][0][:3]],
            "leaf0_head": [int(x) for x in res["leaves"][0][:3]],
            "leaf_last_tail": int(res["leaves"][N - 1][-1]),
            "digests_xor": [int(x) for x in np.bitwise_xor.reduce(res["digests"], axis=0)] if res["digests"].size else [0, 0, 0, 0],
        })
    return {"_comment": "oracle-derived (oracle/oracle.c); the reference pins nothing here — parity unpinned", "cases": cases}


def proof_words(proof):
    """Every field element of a proof in a fixed order (caps, final_poly, witness, query answers)."""
    words = []
    for cap in proof["caps"]:
        words += [int(x) for x in cap.reshape(-1)]
    for a, b in proof["final_poly"]:
        words += [int(a), int(b)]
    words.append(int(proof["pow_witness"]))
    for rnd in proof["rounds"]:
        for row, sib in rnd["initial"]:
            words += [int(x) for x in row] + [int(x) for x in sib.reshape(-1)]
        for ev, sib in rnd["steps"]:
            words += [int(x) for x in ev.reshape(-1)] + [int(x) for x in sib.reshape(-1)]
    return words


def fnv(words):
    d = 0xcbf29ce484222325
    for w in words:
        d = ((d ^ w) * 0x100000001b3) % 2**64
    return d


def opening_proof_vectors():
    import numpy as np
    from oracle import oracle as O
    from oracle import fri_ref as F
    cases = []
    for (n_log, ks, r, h, arities, pow_bits, nq, mul_by_x) in [
            (5, (3, 2), 2, 1, (2, 1), 6, 4, True), (5, (3, 2), 2, 1, (2, 1), 6, 4, False),
            (6, (4, 1, 2), 3, 2, (3, 2), 8, 5, True), (4, (2,), 3, 0, (), 4, 3, True), (8, (5, 3), 3, 4, (4,), 10, 6, True)]:
        n = 1 << n_log
        commits = [O.commit(O.synthetic_values(k, n, seed=o + 1), r, h, is_coeffs=True) for o, k in enumerate(ks)]
        zeta = (int(O.splitmix64(np.array([1000 + n_log], dtype=np.uint64))[0]) % O.P,
                int(O.splitmix64(np.array([2000 + n_log], dtype=np.uint64))[0]) % O.P)
        gzeta = F.escale(zeta, F.root(n_log))
        batches = [(zeta, [(o, i) for o, k in enumerate(ks) for i in range(k)]), (gzeta, [(len(ks) - 1, 0)])]
        ch = F.Challenger()
        for c in commits:
            ch.observe_cap(c["cap"])
        proof = F.prove_openings(commits, batches, ch, r, h, arities, pow_bits, nq, mul_by_x)
        openings = [[tuple(int(v) for v in O.eval_ext2(commits[o]["coeffs"][i], np.array(pt, dtype=np.uint64))) for o, i in polys]
                    for pt, polys in batches]
        fresh = F.Challenger()
        for c in commits:
            fresh.observe_cap(c["cap"])
        assert F.verify(proof, [c["cap"] for c in commits], batches, openings, fresh, n_log, r, h, arities, pow_bits, nq, mul_by_x)
        cases.append({
            "n_log": n_log, "ks": list(ks), "rate_bits": r, "cap_height": h, "arities": list(arities), "pow_bits": pow_bits,
            "num_queries": nq, "mul_by_x": mul_by_x,
            "input": "oracle o = from_coeffs(splitmix64(((o+1) << 48) + c*n + i) mod p); transcript = observe_cap of each oracle",
            "zeta": list(zeta), "points": "batch 0: every polynomial at zeta; batch 1: polynomial 0 of the last oracle at w_n * zeta",
            "final_poly": [[int(a), int(b)] for a, b in proof["final_poly"]],
            "commit_phase_caps_row0": [[int(x) for x in cap[0]] for cap in proof["caps"]],
            "pow_witness": int(proof["pow_witness"]),
            "x_indices": [int(rnd["x_index"]) for rnd in proof["rounds"]],
            "proof_fnv1a64": "%016x" % fnv(proof_words(proof)),
        })
    return {"_comment": "oracle-derived (oracle/fri_ref.py), accepted by its verifier; the reference pins nothing here — parity unpinned",
            "cases": cases}


if __name__ == "__main__":
    with open(os.path.join(HERE, "reference_poseidon_kats.json"), "w") as f:
        json.dump(reference_kats(), f, indent=1)
    with open(os.path.join(HERE, "commit_vectors.json"), "w") as f:
        json.dump(commit_vectors(), f, indent=1)
    with open(os.path.join(HERE, "opening_proof_vectors.json"), "w") as f:
        json.dump(opening_proof_vectors(), f, indent=1)
    print("wrote reference_poseidon_kats.json, commit_vectors.json, opening_proof_vectors.json")
