#!/usr/bin/env bash
# make_reference_fixture.sh — produce REAL zokrates_ark fixtures that pin this repo's proof values to the reference.
#
# The authoring container has no Rust toolchain and the arkworks crates are not vendored, so proof VALUES are "parity
# unpinned" until this script has been run once on any machine with cargo + network access to crates.io and the output
# committed under tests/golden/ref_<name>/ (a few hundred kB per program; sha256 ~30 MB of proving.key is stored
# xz-compressed).  tests/test_reference_fixture.py consumes the fixtures when present (and x-fails, naming this script,
# when absent):
#   * the CPU oracle must reproduce proof.json bit for bit from (out, witness, proving.key, entropy)   [-m "not gpu"]
#   * the GPU path (tools/zkb_generate_proof.py through the C ABI) must write the same proof.json      [-m gpu]
#
# usage: tools/make_reference_fixture.sh /path/to/ZoKrates [outdir = tests/golden]
# Follows the reference's own CLI flow (zokrates_cli/src/ops/{compile,setup,compute_witness,generate_proof}.rs):
#   zokrates compile -i prog.zok -o out            (--stdlib-path <ref>/zokrates_stdlib/stdlib)
#   zokrates setup -i out -b ark -s g16 -e <setup entropy> -p proving.key -v verification.key
#   zokrates compute-witness -i out -o witness -a <args>
#   zokrates generate-proof -i out -w witness -p proving.key -b ark -s g16 -e <proof entropy> -j proof.json
set -euo pipefail
REF=${1:?path to the ZoKrates checkout}
OUT=${2:-"$(cd "$(dirname "$0")/.." && pwd)/tests/golden"}
command -v cargo >/dev/null || { echo "cargo not found: run this where a Rust toolchain exists" >&2; exit 2; }
( cd "$REF" && cargo build --release -p zokrates_cli )
ZOK="$REF/target/release/zokrates"
STDLIB="$REF/zokrates_stdlib/stdlib"
REV=$(cd "$REF" && git rev-parse HEAD 2>/dev/null || echo unknown)

make_one() {  # name, source, curve, "args"
  local name=$1 src=$2 curve=$3 args=$4
  local dir="$OUT/ref_${name}_${curve}"
  mkdir -p "$dir"
  ( cd "$dir"
    "$ZOK" compile -i "$src" -o out -c "$curve" --stdlib-path "$STDLIB"
    "$ZOK" setup -i out -b ark -s g16 -e "fixture-setup-$name" -p proving.key -v verification.key
    # shellcheck disable=SC2086
    "$ZOK" compute-witness -i out -o witness -a $args
    "$ZOK" generate-proof -i out -w witness -p proving.key -b ark -s g16 -e "fixture-proof-$name" -j proof.json
    "$ZOK" verify -j proof.json -v verification.key -b ark
    rm -f abi.json out.r1cs out.wtns
    xz -9 -f proving.key                     # tests read proving.key.xz
    cat > meta.json <<META
{"program": "$(basename "$src")", "curve": "$curve", "arguments": "$args", "setup_entropy": "fixture-setup-$name",
 "proof_entropy": "fixture-proof-$name", "backend": "ark", "scheme": "g16", "zokrates_rev": "$REV",
 "zokrates_version": "$("$ZOK" --version | head -1)"}
META
  )
  echo "wrote $dir"
}

make_one factorize "$REF/zokrates_cli/examples/book/factorize.zok" bn128 "337 113569"
make_one factorize "$REF/zokrates_cli/examples/book/factorize.zok" bls12_381 "337 113569"
make_one sha256 "$REF/zokrates_cli/examples/book/sha256_tutorial/hashexample_updated.zok" bn128 "0 0 0 5"
echo "commit tests/golden/ref_*/ — tests/test_reference_fixture.py picks them up"
