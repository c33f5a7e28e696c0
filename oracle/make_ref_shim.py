"""Build a Python-3 importable *scratch* copy of the reference under ``oracle/_ref/``.

TEST INFRASTRUCTURE ONLY.  The reference (perrysou/SoftGNSS-python) is Python 2.7
source and cannot be imported by the interpreter in this image.  This script
reads the read-only tree at ``/root/reference`` (present only in the build
container, never on the GPU box), applies the mechanical edits listed in
SURVEY.md Appendix B *in memory* and writes the result to ``oracle/_ref/``, which
is git-ignored: no reference source ever enters the history.  The shimmed copy
is used for exactly two things:

* ``tests/golden/make_golden.py`` runs it to produce the committed golden
  input/output vectors that pin ``oracle/gnss_oracle.py``;
* ``tests/test_oracle_vs_reference.py`` (skipped when ``/root/reference`` is
  absent) diffs the oracle restatement against it on fresh inputs.

Edits (all syntactic, none change arithmetic):
  print statements -> print() ; np.long/np.int/long -> int ; np.Inf -> np.inf ;
  np.core.records -> np.rec ; fid.seek(float) -> fid.seek(int(float)) ;
  map() sliced as list ; py2 integer '/' -> '//' at postNavigation.py:584.
"""
import os
import re
import sys
import warnings

REF = os.environ.get("SGX_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
MODULES = ["initialize.py", "acquisition.py", "tracking.py", "postNavigation.py",
           "ephemeris.py", os.path.join("geoFunctions", "__init__.py")]


def _depth_and_continuation(line, depth, in_str):
    """Scan one physical line; return (paren depth, open string delimiter, ends_with_backslash)."""
    i, n = 0, len(line)
    while i < n:
        ch = line[i]
        if in_str:
            if ch == "\\":
                i += 2
                continue
            if line.startswith(in_str, i):
                i += len(in_str)
                in_str = None
                continue
        else:
            if ch == "#":
                break
            if line.startswith('"""', i) or line.startswith("'''", i):
                in_str = line[i:i + 3]
                i += 3
                continue
            if ch in "'\"":
                in_str = ch
            elif ch in "([{":
                depth += 1
            elif ch in ")]}":
                depth -= 1
        i += 1
    if in_str in ("'", '"'):
        in_str = None  # single-quoted strings cannot span lines
    return depth, in_str, line.rstrip().endswith("\\")


def _fix_prints(src):
    lines = src.split("\n")
    out = []
    i = 0
    pat = re.compile(r"^(\s*)print\s+(.*)$")
    while i < len(lines):
        m = pat.match(lines[i])
        if not m or lines[i].lstrip().startswith("#"):
            out.append(lines[i])
            i += 1
            continue
        indent, rest = m.group(1), m.group(2)
        body = [rest]
        depth, in_str, cont = _depth_and_continuation(rest, 0, None)
        while (depth > 0 or cont) and i + 1 < len(lines):
            i += 1
            body.append(lines[i])
            depth, in_str, cont = _depth_and_continuation(lines[i], depth, in_str)
        joined = "\n".join(b[:-1].rstrip() if b.rstrip().endswith("\\") else b for b in
                           [x.rstrip() for x in body])
        joined = joined.rstrip()
        if joined.endswith(","):
            joined = joined[:-1] + ", end=' '"
        out.append("%sprint(%s)" % (indent, joined))
        i += 1
    return "\n".join(out)


def transform(name, src):
    src = _fix_prints(src)
    src = re.sub(r"\bnp\.long\(", "int(", src)
    src = re.sub(r"\bnp\.int\(", "int(", src)
    src = re.sub(r"(?<![\w.])long\(", "int(", src)
    src = src.replace("np.Inf", "np.inf")
    src = src.replace("np.core.records", "np.rec")
    if name == "tracking.py":
        src = src.replace(
            "fid.seek(settings.skipNumberOfBytes + channel[channelNr].codePhase, 0)",
            "fid.seek(int(settings.skipNumberOfBytes + channel[channelNr].codePhase), 0)")
    if name == "postNavigation.py":
        src = re.sub(r"=\s*map\(", "= list(map(", src)
        # close the extra paren opened above on the same logical line
        fixed = []
        for ln in src.split("\n"):
            if "= list(map(" in ln:
                ln = ln.rstrip() + ")"
            fixed.append(ln)
        src = "\n".join(fixed)
        src = src.replace("(len(tlmXcorrResult) + 1) / 2", "(len(tlmXcorrResult) + 1) // 2")
    return src


def build(out_dir=OUT, verbose=False):
    if not os.path.isdir(REF):
        # away from the build container (GPU box): a scratch copy generated earlier may have travelled with the snapshot
        if all(os.path.exists(os.path.join(out_dir, m)) for m in MODULES):
            return out_dir
        raise FileNotFoundError("reference tree %s not present (expected on the GPU box)" % REF)
    os.makedirs(os.path.join(out_dir, "geoFunctions"), exist_ok=True)
    for rel in MODULES:
        with open(os.path.join(REF, rel), "r") as f:
            src = f.read()
        new = transform(os.path.basename(rel) if "geoFunctions" not in rel else rel, src)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")  # LaTeX plot labels trip SyntaxWarning
            compile(new, rel, "exec")  # self-check: must parse under this interpreter
        with open(os.path.join(out_dir, rel), "w") as f:
            f.write(new)
        if verbose:
            print("shimmed", rel)
    return out_dir


def import_ref():
    """Return the shimmed reference modules (initialize, acquisition, tracking, postNavigation)."""
    path = build()
    if path not in sys.path:
        sys.path.insert(0, path)
    import importlib
    warnings.filterwarnings("ignore", category=SyntaxWarning)
    mods = {}
    for m in ("initialize", "acquisition", "tracking", "postNavigation", "ephemeris"):
        mods[m] = importlib.import_module(m)
    return mods


if __name__ == "__main__":
    build(verbose=True)
