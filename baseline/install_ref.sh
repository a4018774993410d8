#!/bin/bash
# Installs the UNMODIFIED reference into baseline/_ref (git-ignored, travels to the GPU box with the gpurun snapshot) so
# that `bench.py --impl reference` can run the reference's own OnPolicyRunner.learn on the box's host cores.
#   1. pip install (offline) of the reference and of its bundled rsl_rl, from a /tmp copy (/root/reference is read-only);
#   2. upstream's setup.py uses find_packages(), which skips every directory without an __init__.py -- legged_gym/envs/base,
#      envs/go2, envs/anymal_c, envs/cassie, scripts: the installed tree cannot even import legged_gym.envs.  Those module
#      directories are completed from the same source tree, untouched (cp -n: nothing pip installed is overwritten).
# Authoring container only; a no-op when /root/reference is absent.
set -e
REF=${1:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
[ -d "$REF/legged_gym" ] || { echo "install_ref: $REF not present, nothing to do"; exit 0; }
TMP=$(mktemp -d)
cp -r "$REF" "$TMP/ref"
rm -rf "$HERE/_ref"; mkdir -p "$HERE/_ref"
python -m pip install -q --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$HERE/_ref" "$TMP/ref"
python -m pip install -q --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$HERE/_ref" "$TMP/ref/rsl_rl"
(cd "$REF" && find legged_gym -name '*.py' -not -path '*(broken)*' -not -path '*/tests/*') | while read -r f; do
  mkdir -p "$HERE/_ref/$(dirname "$f")"
  cp --update=none "$REF/$f" "$HERE/_ref/$f"
done
rm -rf "$TMP"
find "$HERE/_ref" -name __pycache__ -prune -exec rm -rf {} +
echo "install_ref: $(find "$HERE/_ref" -name '*.py' | wc -l) python files under baseline/_ref"
