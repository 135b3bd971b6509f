"""CPU oracle for the Groth16 proving hot path — TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement of what `zokrates_ark` + arkworks 0.3.0 compute on the
`generate-proof` path.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import it.  The product
(`zokrates_b200/`) never does.

Parity status:
  * field arithmetic, byte formats, witness semantics: PINNED against the reference's own
    KATs (zokrates_field/src/bn128.rs tests, ir/witness.rs tests, sha256 512bitPacked KAT,
    curve ids from zokrates_book/src/toolbox/ir.md).
  * proof VALUES (MSM / NTT / group results): "parity unpinned" — the reference holds no
    golden proof / proving key and cannot be compiled here (no rustc,
    arkworks sources absent).  They are validated algebraically instead (pairing equation,
    trapdoor check, two independent implementations: python big-int and C).
"""
