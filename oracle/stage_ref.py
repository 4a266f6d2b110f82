"""TEST INFRASTRUCTURE -- stage the UNMODIFIED reference for machines that do not have /root/reference.

    python -m oracle.stage_ref          (also run by __graft_entry__.build() when /root/reference exists)

The reference is eight flat Python files with nothing to compile, so the "build" of `oracle/_ref/` is an archive:
`oracle/_ref/dreamgnn_reference.zip` holds the `.py` files byte for byte (Python imports modules straight from a zip on
sys.path) plus `MANIFEST.json` with their SHA-256. `oracle/_ref/` is git-ignored -- reference sources never enter the
history -- but NOT gpurun-ignored, so the archive travels to the GPU box with the snapshot. There it is what
`bench.py --impl reference` times (the reference's own `train.py` step on the box's host cores, through the DGL stand-in
in oracle/dgl) and what `tests/test_gpu_dropin.py` imports to run the reference's own `model.py` / `train.py` on top of
this repo's `layers` and graph handle. Nothing in the product package reads it.
"""
import hashlib
import json
import os
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_DIR = os.environ.get('DREAMGNN_REFERENCE_DIR', '/root/reference')
REF_DIR = os.path.join(HERE, '_ref')
ARCHIVE = os.path.join(REF_DIR, 'dreamgnn_reference.zip')
MODULES = ('utils', 'augmentation', 'data_loader', 'layers', 'model', 'evaluation', 'train')


def stage(force=False):
    """Write the archive (idempotent). Returns its path, or None when the reference tree is not present."""
    if not os.path.isfile(os.path.join(REFERENCE_DIR, 'layers.py')):
        return ARCHIVE if os.path.isfile(ARCHIVE) else None
    manifest = {}
    for m in MODULES:
        with open(os.path.join(REFERENCE_DIR, m + '.py'), 'rb') as fh:
            manifest[m + '.py'] = hashlib.sha256(fh.read()).hexdigest()
    if not force and os.path.isfile(ARCHIVE):
        try:
            with zipfile.ZipFile(ARCHIVE) as z:
                if json.loads(z.read('MANIFEST.json')) == manifest:
                    return ARCHIVE
        except (KeyError, ValueError, zipfile.BadZipFile):
            pass
    os.makedirs(REF_DIR, exist_ok=True)
    tmp = ARCHIVE + '.tmp'
    with zipfile.ZipFile(tmp, 'w', zipfile.ZIP_DEFLATED) as z:
        for m in MODULES:
            z.write(os.path.join(REFERENCE_DIR, m + '.py'), m + '.py')
        z.writestr('MANIFEST.json', json.dumps(manifest, indent=1, sort_keys=True))
    os.replace(tmp, ARCHIVE)
    return ARCHIVE


def available():
    return os.path.isfile(ARCHIVE)


if __name__ == '__main__':
    p = stage(force='--force' in sys.argv)
    print(p or 'reference tree not present at %s and no staged archive' % REFERENCE_DIR)
