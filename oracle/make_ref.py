#!/usr/bin/env python
"""Stage the UNMODIFIED reference implementation next to the oracle:  /root/reference/cliora -> oracle/_ref/cliora.

TEST / BASELINE INFRASTRUCTURE ONLY.  The reference is pure Python (no build step), so "building" it is a byte-for-
byte copy of its package directory.  ``oracle/_ref/`` is git-ignored (reference sources never enter this repo's
history) but not gpurun-ignored, so the staged copy travels to the GPU box, where ``bench.py --impl reference`` and
``bench.py``'s ``cpu_baseline`` / ``gpu_baseline`` legs import it and drive its own ``build_net`` + ``Trainer.step``
(cliora/net/trainer.py:483-582) -- ``kind: "reference"``.  Nothing under ``cliora_b200/`` may import it.

    python oracle/make_ref.py            # in the build container (where /root/reference exists)

``__graft_entry__.build()`` runs this when /root/reference is present.  A manifest with the sha256 of every copied
file is written to oracle/_ref/MANIFEST.json so the copy can be checked against the checkout it came from.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get('CLIORA_REFERENCE', '/root/reference')
DST = os.path.join(HERE, '_ref')


def stage(src=SRC, dst=DST):
    pkg = os.path.join(src, 'cliora')
    if not os.path.isdir(pkg):
        return False
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(dst)
    shutil.copytree(pkg, os.path.join(dst, 'cliora'), ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
    manifest = {}
    for root, _, files in os.walk(os.path.join(dst, 'cliora')):
        for f in sorted(files):
            p = os.path.join(root, f)
            with open(p, 'rb') as fh:
                manifest[os.path.relpath(p, dst)] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(dst, 'MANIFEST.json'), 'w') as fh:
        json.dump({'source': src, 'files': manifest}, fh, indent=1, sort_keys=True)
    return True


if __name__ == '__main__':
    ok = stage()
    print('staged %s -> %s' % (SRC, DST) if ok else 'no reference checkout at %s; nothing staged' % SRC)
    sys.exit(0)
