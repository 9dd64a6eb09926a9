#!/bin/bash
# Build a compile-time view of the (read-only) reference tree: real directories, every file a symlink
# into $REF, except Lights/Light.h which is regenerated through sed because g++ rejects
# `static [[nodiscard]] auto` (Lights/Light.h:280,302), plus a `Shapes -> shapes` alias for the
# wrong-case include in base/STLReader.cpp:4.  No reference source is copied into git: the output
# directory is ignored and is rebuilt from /root/reference whenever needed.
set -euo pipefail
REF=${1:?reference root}
OUT=${2:?output dir}
rm -rf "$OUT"
mkdir -p "$OUT"
(cd "$REF" && find . -type d -not -path './.git*') | while read -r d; do mkdir -p "$OUT/$d"; done
(cd "$REF" && find . -type f -not -path './.git/*' \( -name '*.h' -o -name '*.cpp' \)) | while read -r f; do
  ln -s "$REF/$f" "$OUT/$f"
done
rm -f "$OUT/Lights/Light.h"
sed 's/static \[\[nodiscard\]\] auto/[[nodiscard]] static auto/' "$REF/Lights/Light.h" > "$OUT/Lights/Light.h"
ln -s shapes "$OUT/Shapes"
touch "$OUT/.stamp"
