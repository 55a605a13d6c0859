"""Install the UNMODIFIED reference (2gunsu/monocon-pytorch) into baseline/_ref/ for `bench.py --impl reference`.

baseline/_ref/ is git-ignored (the reference's sources never enter this repository's history) but NOT gpurun-ignored, so it
travels to the GPU box with the snapshot.  Runs in the build container only (where /root/reference exists):

  1. the prescribed offline install
        python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /root/reference
     -- the reference has neither setup.py nor pyproject.toml, so pip refuses it ("does not appear to be a Python project");
  2. therefore the documented fallback of BASELINE.md section 4: a verbatim copy of the reference tree (Python packages and
     entry scripts; `resources/` images left out).
Idempotent; prints what it did."""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = '/root/reference'
DST = os.path.join(ROOT, 'baseline', '_ref')


def main() -> int:
    if not os.path.isdir(SRC):
        print(f'{SRC} is absent (GPU box): using the prebuilt {DST}' if os.path.isdir(DST) else f'{SRC} is absent and {DST} was never installed')
        return 0
    if os.path.exists(os.path.join(DST, 'model', '__init__.py')):
        print(f'{DST} already installed')
        return 0
    os.makedirs(os.path.dirname(DST), exist_ok=True)
    r = subprocess.run([sys.executable, '-m', 'pip', 'install', '--no-index', '--no-build-isolation', '--find-links', '/opt/wheelhouse',
                        '--target', DST, SRC], capture_output=True, text=True)
    if r.returncode == 0 and os.path.exists(os.path.join(DST, 'model', '__init__.py')):
        print('pip install succeeded')
        return 0
    print('pip install refused the reference (no setup.py / pyproject.toml): ' + (r.stderr.strip().splitlines() or ['?'])[-1])
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns('resources', '.git', '__pycache__', '*.pyc'))
    with open(os.path.join(DST, 'INSTALLED_FROM.txt'), 'w') as f:
        f.write(f'verbatim copy of {SRC} (unmodified reference) made by scripts/install_reference.py; git-ignored\n')
    print(f'copied the reference tree to {DST}')
    return 0


if __name__ == '__main__':
    sys.exit(main())
