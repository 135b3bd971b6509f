#!/usr/bin/env python3
"""Ingestion and CLI-level wall time at benchmark size (SURVEY.md §8d "like-for-like CLI wall time").

Writes the three files `zokrates generate-proof` reads — the compiled program `out`, the binary `witness`, `proving.key` — for
the synthetic 2^k circuit of bench.py, then times, on one GPU:
  * the native front door: zkb_prog_load (parse + ark-order synthesis + level schedule + upload), zkb_prog_compute_witness
    (device interpreter; checked against the generator's assignment), zkb_prog_set_witness (witness file -> assignment);
  * `B200.generate_proof_files` (the trait-shaped call: program + witness + key bytes -> proof) cold and again with the same
    key bytes (content-hash key cache: the window tables are not rebuilt);
  * the file-level tool in a FRESH process (`tools/zkb_generate_proof.py`, cold CUDA context, key read from disk, no tables);
  * for scale, the Python reader of the same file format on a 2^14 sample (the round-1 path).
Prints one JSON object; run under gpurun.  Not a bench.py line.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from multiprocessing import Pool

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

_G = {}


def _head(major, n):
    if n < 24:
        return bytes([major << 5 | n])
    if n < 1 << 8:
        return bytes([major << 5 | 24, n])
    if n < 1 << 16:
        return bytes([major << 5 | 25]) + n.to_bytes(2, "big")
    if n < 1 << 32:
        return bytes([major << 5 | 26]) + n.to_bytes(4, "big")
    return bytes([major << 5 | 27]) + n.to_bytes(8, "big")


_LC = b"\xa2\x64span\xf6\x65value"
_ST0 = b"\xa1\x6aConstraint\xa4\x64span\xf6\x64quad\xa3\x64span\xf6\x64left"
_ID = b"\x82\xa1\x62id"


def _chunk(args):
    lo, hi = args
    out = []
    mats = _G["mats"]
    for i in range(lo, hi):
        parts = [_ST0]
        for k, key in enumerate((None, b"\x65right", b"\x63lin")):
            rp, col, val = mats[k]
            a, b = int(rp[i]), int(rp[i + 1])
            if key:
                parts.append(key)
            parts.append(_LC + _head(4, b - a))
            for t in range(a, b):
                parts.append(_ID + _head(0, int(col[t])) + b"\x58\x20" + val[t].tobytes())
        parts.append(b"\x65error\xf6")
        out.append(b"".join(parts))
    return b"".join(out)


def write_out_file(r1cs, curve_name, n_public, n_private, procs):
    """The synthetic circuit as a compiled-program file: variable id = matrix column (the generator numbers its columns in ark's
    allocation order already), arguments = the instance / private-input columns, one Constraint statement per row."""
    from zokrates_b200 import zir
    import io
    import struct
    _G["mats"] = [(np.asarray(m[0]), np.asarray(m[1]), np.ascontiguousarray(m[2], dtype="<u8")) for m in r1cs.matrices()]
    N = r1cs.num_constraints
    step = max(1, N // (procs * 4))
    spans = [(a, min(N, a + step)) for a in range(0, N, step)]
    if procs > 1:
        with Pool(procs) as pool:                      # fork: the matrices are inherited, only byte strings come back
            body_parts = pool.map(_chunk, spans)
    else:
        body_parts = [_chunk(s) for s in spans]
    body = io.BytesIO()
    body.write(b"\x00" * zir.HEADER_RESERVED)
    sec = []
    a = body.tell()
    params = [{"span": None, "id": {"id": 1 + i}, "private": i >= n_public} for i in range(n_public + n_private)]
    zir.cbor_encode(params, body); sec.append((a, body.tell() - a))
    a = body.tell()
    for p in body_parts:
        body.write(p)
    sec.append((a, body.tell() - a))
    a = body.tell(); zir.cbor_encode([], body); sec.append((a, body.tell() - a))
    a = body.tell(); zir.cbor_encode({"modules": {}}, body); sec.append((a, body.tell() - a))
    head = zir.MAGIC + zir.VERSION + zir.CURVE_IDS[curve_name] + struct.pack("<II", N, 0)
    for ty, (off, ln) in zip(zir.SECTION_TYPES, sec):
        head += struct.pack("<IQQ", ty, off, ln)
    buf = bytearray(body.getvalue())
    buf[:len(head)] = head
    return bytes(buf)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=20)
    ap.add_argument("--curve", default="bn128")
    ap.add_argument("--procs", type=int, default=min(32, os.cpu_count() or 1))
    ap.add_argument("--skip-cli", action="store_true")
    args = ap.parse_args()
    from zokrates_b200 import backend, rng, synthetic, zir
    from zokrates_b200._lib import Context, Library, fr_from_array, OPT_TABLES
    from zokrates_b200.curves import curve as get_curve
    c = get_curve(args.curve)
    ctx = backend.context(c, 0)
    n_cons = (1 << args.log_n) - 2
    res = {"curve": args.curve, "constraints": n_cons, "host_cpus": os.cpu_count()}
    r1cs, z = synthetic.make_layered(ctx, args.curve, n_cons)
    t = time.perf_counter()
    out_bytes = write_out_file(r1cs, args.curve, 1, 6, args.procs)
    res["write_out_file_s"] = round(time.perf_counter() - t, 2)
    res["out_bytes"] = len(out_bytes)

    # -- native front door
    t = time.perf_counter(); prog = ctx.prog_load(out_bytes); res["prog_load_s"] = round(time.perf_counter() - t, 3)
    info = ctx.prog_info(prog)
    res["prog_info"] = {k: info[k] for k in ("constraints", "instance", "witness", "levels", "directives")}
    inputs = fr_from_array(z[1:8])
    t = time.perf_counter(); wit = ctx.prog_compute_witness(prog, inputs); res["compute_witness_s"] = round(time.perf_counter() - t, 3)
    res["compute_witness_device_ms"] = ctx.timings().get("witness_eval")
    rec = np.frombuffer(wit, dtype=np.uint8)[8:].reshape(-1, 40)
    got = np.ascontiguousarray(rec[:, 8:]).view("<u8").reshape(-1, 4)
    assert len(got) == len(z) and np.array_equal(got, z), "device interpreter disagrees with the generator's assignment"
    res["witness_bytes"] = len(wit)
    t = time.perf_counter(); ctx.prog_set_witness(prog, wit); res["set_witness_s"] = round(time.perf_counter() - t, 3)
    t = time.perf_counter(); pk = ctx.setup(info["r1cs"], [11, 22, 33, 44, 5555, 3, 7]); res["setup_s"] = round(time.perf_counter() - t, 2)
    res["pk_bytes"] = len(pk)
    ctx.prog_free(prog)

    # -- the trait-shaped call, three times with the same bytes: cold key, then the content-hash cache
    calls = []
    for k in range(3):
        tm = {}
        t = time.perf_counter()
        proof = backend.B200.generate_proof_files(out_bytes, wit, pk, rng.get_rng_from_entropy("ingest"), curve=args.curve, timings=tm)
        tm["total_s"] = time.perf_counter() - t
        calls.append({k2: round(v, 4) for k2, v in tm.items()})
    res["generate_proof_files_calls"] = calls
    res["proof_a_x"] = proof.to_tagged_json()[:120].split('"a"')[-1][:90] if proof else None

    # -- the Python reader on a small sample of the same family, for scale
    r_s, z_s = synthetic.make_layered(ctx, args.curve, (1 << 14) - 2)
    small = write_out_file(r_s, args.curve, 1, 6, 1)
    t = time.perf_counter(); zir.read_prog(small); res["python_read_prog_2p14_s"] = round(time.perf_counter() - t, 2)
    t = time.perf_counter(); h = ctx.prog_load(small); ctx.prog_free(h); res["native_prog_load_2p14_s"] = round(time.perf_counter() - t, 4)

    # -- the file-level tool in a fresh process: what a user of `zokrates generate-proof` waits for
    if not args.skip_cli:
        with tempfile.TemporaryDirectory() as d:
            for name, data in (("out", out_bytes), ("witness", wit), ("proving.key", pk)):
                with open(os.path.join(d, name), "wb") as f:
                    f.write(data)
            cmd = [sys.executable, os.path.join(ROOT, "tools", "zkb_generate_proof.py"), "-i", os.path.join(d, "out"), "-w", os.path.join(d, "witness"),
                   "-p", os.path.join(d, "proving.key"), "-j", os.path.join(d, "proof.json"), "-e", "ingest", "--verbose"]
            runs = []
            for _ in range(2):
                t = time.perf_counter()
                p = subprocess.run(cmd, capture_output=True, text=True)
                wall = time.perf_counter() - t
                tl = [l for l in p.stdout.splitlines() if l.startswith("timings:")]
                runs.append({"wall_s": round(wall, 3), "rc": p.returncode, "timings": tl[-1] if tl else p.stderr[-300:]})
            res["cli_generate_proof"] = runs
            same = json.load(open(os.path.join(d, "proof.json")))["proof"] == json.loads(proof.to_tagged_json())["proof"]
            res["cli_proof_equals_resident_proof"] = bool(same)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
