#!/usr/bin/env bash
# Guarded timing of faer's OWN rayon CPU path at this repo's shapes. It needs cargo + rustc and the
# crates faer depends on; the build image has neither (no network, no vendored registry), so there this prints one
# {"impl": "faer-rayon", "unavailable": ...} line and exits 0. On a box with a Rust toolchain:
#   FAER_SRC=/path/to/faer-rs/faer tools/faer_rayon_bench.sh [gemm|llt|lu|qr|svd|all] [n]
set -u
here="$(cd "$(dirname "$0")" && pwd)"
if ! command -v cargo >/dev/null 2>&1 || ! command -v rustc >/dev/null 2>&1; then
  echo '{"impl": "faer-rayon", "unavailable": "no Rust toolchain (cargo / rustc) on this machine"}'
  exit 0
fi
src="${FAER_SRC:-/root/reference/faer}"
if [ ! -f "$src/Cargo.toml" ]; then
  echo "{\"impl\": \"faer-rayon\", \"unavailable\": \"faer sources not found at $src (set FAER_SRC)\"}"
  exit 0
fi
work="$(mktemp -d)"
cp -r "$here/faer_rayon_bench/." "$work/"
sed -i "s#path = \"../../../reference/faer\"#path = \"$src\"#" "$work/Cargo.toml"
if ! (cd "$work" && cargo build --release --quiet 2> "$work/build.log"); then
  echo "{\"impl\": \"faer-rayon\", \"unavailable\": \"cargo build failed: $(tail -1 "$work/build.log" | tr -d '\"')\"}"
  exit 0
fi
"$work/target/release/faer_rayon_bench" "${1:-all}" ${2:-}
