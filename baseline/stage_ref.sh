#!/bin/bash
# Stage the UNMODIFIED reference (ddbourgin/numpy-ml v0.1.2, pure Python) under the git-ignored baseline/_ref/ so that
# it travels to the GPU box with the repository snapshot (it is git-ignored, not gpurun-ignored).  Run in the build
# container, where /root/reference exists; --no-deps because numpy / scipy are already in the image and the offline
# wheelhouse has no numpy wheel to "resolve".
set -e
cd "$(dirname "$0")/.."
rm -rf baseline/_ref
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref /root/reference
ls baseline/_ref
