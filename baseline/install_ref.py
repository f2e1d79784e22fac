"""Install the UNMODIFIED reference (utiasASRL/dICP) into git-ignored `baseline/_ref/`.

    python baseline/install_ref.py            # no-op when /root/reference is absent (GPU box)

This is the base contract's one offline install: `pip install --no-index --no-build-isolation
--no-deps --target baseline/_ref <copy of /root/reference>` (from a copy under /tmp, because the
build writes `build/` and `*.egg-info` into the source tree and /root/reference is read-only;
`--no-deps` because its only undeclared import that is missing here, matplotlib, is unused on the
compute path and stubbed by `baseline/ref_arm.py`).  The wheel holds the package `dICP/` only, so
the default configuration the package looks for at `<package>/../config/dICP_config.yaml`
(`ICP.py:17-19`) is copied next to it.  `baseline/_ref/` is listed in `.gitignore` (no reference
source enters the history) but not in `.gpurunignore`, so it travels to the GPU box, where
`bench.py --impl reference` and the `cpu_baseline` leg time it.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("DICP_REFERENCE_SRC", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")


def installed():
    return os.path.exists(os.path.join(REF_DST, "dICP", "ICP.py")) and \
        os.path.exists(os.path.join(REF_DST, "config", "dICP_config.yaml"))


def install(force=False):
    """Returns 'installed', 'present' or 'unavailable: <why>'."""
    if installed() and not force:
        return "present"
    if not os.path.isdir(REF_SRC):
        return "unavailable: %s does not exist (the GPU box only carries a prebuilt baseline/_ref)" % REF_SRC
    tmp = tempfile.mkdtemp(prefix="dicp_ref_")
    try:
        work = os.path.join(tmp, "reference")
        shutil.copytree(REF_SRC, work)
        if os.path.isdir(REF_DST):
            shutil.rmtree(REF_DST)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
               "--find-links", "/opt/wheelhouse", "--target", REF_DST, work]
        proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if proc.returncode != 0:
            return "unavailable: pip install failed: " + proc.stdout.strip().splitlines()[-1]
        shutil.copytree(os.path.join(REF_SRC, "config"), os.path.join(REF_DST, "config"))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return "installed" if installed() else "unavailable: install produced no dICP/ICP.py"


if __name__ == "__main__":
    print("baseline/_ref:", install(force="-f" in sys.argv))
