#!/bin/bash
# Installs the UNMODIFIED upstream reference into baseline/_ref (git-ignored).
# The upstream tree has no setup.py/pyproject, so `pip install /root/reference` fails with
# "Neither 'setup.py' nor 'pyproject.toml' found"; we install from a /tmp copy that only
# adds the packaging shim in baseline/ref_packaging/setup.py (no source edits).
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
SRC=${1:-/root/reference}
rm -rf /tmp/mine_ref && cp -r "$SRC" /tmp/mine_ref
cp "$HERE/ref_packaging/setup.py" /tmp/mine_ref/setup.py
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$HERE/_ref" --upgrade /tmp/mine_ref
