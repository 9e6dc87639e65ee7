"""Copy the few reference files the CPU arm of bench.py executes into the git-ignored baseline/_ref/.

TEST / MEASUREMENT INFRASTRUCTURE ONLY.   python oracle/install_reference.py

The reference is script-level Python without a package (no setup.py / pyproject: `pip install /root/reference` does not
apply), and /root/reference does not exist on the GPU box.  baseline/_ref/ is listed in .gitignore (reference sources
never enter this repository's history) but not in .gpurunignore, so the copy travels with the snapshot and
`bench.py --impl reference` can exec the UNMODIFIED LeMoC.py there under oracle/shim (astropy / shapely stand-ins).
Run by __graft_entry__.build() whenever /root/reference is present.
"""
from __future__ import annotations

import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = ("LeMoC/LeMoC.py", "LeMoC/LeHaMoC_f.py", "LeMoC/Parameters_Test1.txt",
         "Fit_emcee/Code.py", "Fit_emcee/LeHaMoC_f.py", "Fit_emcee/Parameters.txt")


def install(force: bool = False) -> bool:
    """-> True when baseline/_ref holds the files afterwards."""
    if not os.path.isdir(os.path.join(SRC, "LeMoC")):
        return all(os.path.exists(os.path.join(DST, f)) for f in FILES)
    for f in FILES:
        src, dst = os.path.join(SRC, f), os.path.join(DST, f)
        if force or not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
    return True


if __name__ == "__main__":
    print("baseline/_ref ready:", install(force=True))
