#!/usr/bin/env bash
# Installs the UNMODIFIED reference into baseline/_ref (git-ignored; travels to the GPU box with the snapshot).
# /root/reference ships no setup.py / pyproject.toml, so pip refuses it directly: the install is
# done from a copy under /tmp to which ONLY a packaging shim (package list, no source change) is added.
set -euo pipefail
REPO="$(cd "$(dirname "$0")/.." && pwd)"
SRC="${1:-/root/reference}"
TMP="$(mktemp -d /tmp/refcopy.XXXXXX)"
trap 'rm -rf "$TMP"' EXIT
cp -r "$SRC/." "$TMP/"
rm -rf "$TMP/.git"
cat > "$TMP/setup.py" <<'EOF'
from setuptools import setup, find_packages
setup(name="fedstil-reference", version="0.0.0",
      packages=find_packages(include=["analyse*", "criterions*", "datasets*", "methods*", "models*", "modules*",
                                      "tools*"]),
      py_modules=["builder", "experiment", "main"])
EOF
# packages without an __init__.py are still directories of modules the reference imports as namespace packages
for d in analyse criterions datasets methods models modules tools; do
  [ -d "$TMP/$d" ] && [ ! -f "$TMP/$d/__init__.py" ] && touch "$TMP/$d/.flpr_namespace" || true
done
rm -rf "$REPO/baseline/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$REPO/baseline/_ref" "$TMP"
# namespace directories (no __init__.py) are not picked up by find_packages: copy them verbatim
for d in analyse criterions datasets methods models modules tools; do
  if [ -d "$TMP/$d" ] && [ ! -d "$REPO/baseline/_ref/$d" ]; then cp -r "$TMP/$d" "$REPO/baseline/_ref/$d"; fi
done
cp -r "$SRC/configs" "$REPO/baseline/_ref/configs"
find "$REPO/baseline/_ref" -name .flpr_namespace -delete
echo "installed reference into $REPO/baseline/_ref"
