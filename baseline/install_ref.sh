#!/bin/bash
# pip needs to write build files next to setup.py and /root/reference is read-only: install from a copy.
set -e
cd "$(dirname "$0")/.."
rm -rf /tmp/rectorch_ref_src baseline/_ref
cp -r /root/reference /tmp/rectorch_ref_src
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target baseline/_ref /tmp/rectorch_ref_src
ls baseline/_ref
