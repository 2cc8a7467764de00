#!/bin/sh
# Installs the UNMODIFIED reference (quanpn90/avsr) into the git-ignored baseline/_ref/ so that `bench.py --impl reference`
# and tests/test_scorers_reference_driver.py can import it on the GPU box (where /root/reference does not exist).
# The reference has no setup.py / pyproject.toml, so `pip install --target baseline/_ref /root/reference` has nothing to
# build ("neither 'setup.py' nor 'pyproject.toml' found"); its packages are plain directories under src/, which is what is
# copied here.  Nothing in baseline/_ref is ever edited or committed (.gitignore), and no product module imports it.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC="${1:-/root/reference}"
[ -d "$SRC/src" ] || { echo "install_ref: $SRC/src not found" >&2; exit 1; }
rm -rf "$ROOT/baseline/_ref"
mkdir -p "$ROOT/baseline/_ref"
cp -r "$SRC/src" "$ROOT/baseline/_ref/src"
find "$ROOT/baseline/_ref" -name __pycache__ -type d -prune -exec rm -rf {} +
( cd "$SRC" && find src -type f -name '*.py' | sort | xargs sha256sum ) > "$ROOT/baseline/_ref/SHA256SUMS"
echo "installed $(find "$ROOT/baseline/_ref/src" -name '*.py' | wc -l) reference files into baseline/_ref"
