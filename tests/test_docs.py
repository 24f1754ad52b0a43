"""Paperwork checks (no GPU): DESIGN.md section 6 is the rendering of the committed measurement files, and every file
under profiles/ / tools/ / tests/ that the documents cite exists."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _read(name):
    with open(os.path.join(ROOT, name)) as f:
        return f.read()


def test_design_section_6_is_rendered_from_profiles():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_measured_md.py")], capture_output=True, text=True, check=True).stdout
    design = _read("DESIGN.md")
    missing = [line for line in out.splitlines() if line.strip() and line not in design]
    assert not missing, "DESIGN.md section 6 is stale; run tools/make_measured_md.py:\n" + "\n".join(missing[:5])


def test_cited_files_exist():
    cited = set()
    for doc in ("DESIGN.md", "INTEGRATION.md", "README.md", os.path.join("profiles", "README.md"), os.path.join("tools", "README.md")):
        text = _read(doc)
        cited |= set(re.findall(r"`((?:profiles|tools|tests|oracle|include|pyqsm_b200|baseline)/[A-Za-z0-9_./-]+\.[a-z]+)`", text))
        if doc.startswith("profiles"):
            cited |= {"profiles/" + m for m in re.findall(r"`(r0[12]_[A-Za-z0-9_.-]+\.(?:txt|json|csv))`", text)}
        if doc.startswith("tools"):
            cited |= {"tools/" + m for m in re.findall(r"`([a-z_0-9]+\.(?:py|sh))`", text) if not os.path.exists(os.path.join(ROOT, m))}
    built = (".so",)                                    # built artefacts are not in the tree
    absent = sorted(p for p in cited if not p.endswith(built) and "*" not in p and not os.path.exists(os.path.join(ROOT, p)))
    assert not absent, absent
