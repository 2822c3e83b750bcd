#!/bin/sh
# Dump known-answer vectors from the reference crate's own arithmetic.
#   tools/ref_kat/run.sh [/path/to/rust-pathtracer checkout] [out.json]
# Needs cargo + network (or a vendored registry) for the reference's dependencies; neither exists in the build image.
set -eu
REF=${1:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=${2:-$HERE/../../tests/golden/ref_kat_f32.json}
WORK=$HERE/_work
rm -rf "$WORK"; mkdir -p "$WORK"
cp -r "$REF/rust-pathtracer" "$WORK/rust-pathtracer"
cp "$REF/renderer/src/analytical.rs" "$WORK/analytical.rs"
T="$WORK/rust-pathtracer/src/tracer.rs"
# 1. private methods of `impl Tracer` (4-space indent; nested helper fns are indented deeper and stay private) become `pub`
sed -i -E 's/^    fn /    pub fn /' "$T"
# 2. the RNG TYPE the drawing functions take becomes the scripted one (same `Rng::gen` call sites, same f32 conversion)
sed -i -E 's/^use rand::\{thread_rng, Rng, rngs::ThreadRng\};/use rand::Rng; use crate::kat_rng::{thread_rng, ThreadRng};/' "$T"
# 3. the scripted RNG module joins the crate
cp "$HERE/kat_rng.rs" "$WORK/rust-pathtracer/src/kat_rng.rs"
printf '\npub mod kat_rng;\n' >> "$WORK/rust-pathtracer/src/lib.rs"
# the complete difference to the reference, for the record (must show nothing but the three edits above)
diff -ru "$REF/rust-pathtracer/src" "$WORK/rust-pathtracer/src" > "$WORK/patch.diff" || true
grep -c '^[-+][^-+]' "$WORK/patch.diff" | sed 's/^/[ref_kat] changed lines vs the reference: /'
(cd "$HERE" && cargo run --release -- "$OUT")
echo "[ref_kat] wrote $OUT — now: python -m pytest tests/test_ref_kat.py -q"
