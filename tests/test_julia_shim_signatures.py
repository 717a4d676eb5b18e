"""Static check of the Julia shim (spacecharge.jl_b200/julia/SpaceChargeB200.jl) against the C header.

Julia is not installed in this image, so the shim cannot be executed here; what can be
checked is that every `ccall` names a symbol the header declares and passes the same number and kind of
arguments (pointer / Int64 / Cint / Float64) in the same order as the C prototype -- a mismatch there is the
error a maintainer would otherwise only see as a crash on the GPU box."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "spacecharge.jl_b200", "julia", "SpaceChargeB200.jl")
HEADERS = [os.path.join(ROOT, "include", h) for h in ("spacecharge_b200.h", "spacecharge_b200_debug.h")]


def split_top(s):
    """Split at commas that are not nested in (), {} or []."""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def c_kind(param):
    p = re.sub(r"/\*.*?\*/", "", param).strip()
    if "*" in p or "[" in p:
        return "ptr"
    if re.match(r"^(const\s+)?int64_t\b", p):
        return "i64"
    if re.match(r"^(const\s+)?int\b", p):
        return "i32"
    if re.match(r"^(const\s+)?double\b", p):
        return "f64"
    raise AssertionError("unclassified C parameter: %r" % param)


def jl_kind(t):
    t = t.strip()
    if t.startswith(("Ptr{", "CuPtr{", "Ref{")) or t == "Cstring":
        return "ptr"
    if t == "Int64":
        return "i64"
    if t == "Cint":
        return "i32"
    if t in ("Float64", "Cdouble"):
        return "f64"
    raise AssertionError("unclassified Julia argument type: %r" % t)


def header_prototypes():
    protos = {}
    for h in HEADERS:
        text = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        for m in re.finditer(r"SCB_API\s+([\w\s\*]+?)\b(scb_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
            params = m.group(3).strip()
            kinds = [] if params in ("", "void") else [c_kind(p) for p in split_top(" ".join(params.split()))]
            protos[m.group(2)] = kinds
    return protos


def matching_paren(text, start):
    depth = 0
    for i in range(start, len(text)):
        if text[i] == "(":
            depth += 1
        elif text[i] == ")":
            depth -= 1
            if depth == 0:
                return i
    raise AssertionError("unbalanced parentheses")


def shim_ccalls():
    text = open(SHIM).read()
    calls = []
    for m in re.finditer(r"ccall\(\(:(scb_\w+), LIB\)", text):
        end = matching_paren(text, m.start() + len("ccall"))
        args = split_top(text[m.start() + len("ccall("):end])
        # args[0] = (:sym, LIB), args[1] = return type, args[2] = tuple of argument types (or a variable bound to one)
        types = args[2]
        if not types.startswith("("):
            bound = re.findall(r"\b%s\s*=\s*\(" % re.escape(types), text[:m.start()])
            assert bound, "type tuple %r of %s is not defined before the call" % (types, m.group(1))
            pos = [b.start() for b in re.finditer(r"\b%s\s*=\s*\(" % re.escape(types), text[:m.start()])][-1]
            open_paren = text.index("(", pos)
            types = text[open_paren:matching_paren(text, open_paren) + 1]
        inner = types.strip()[1:-1]
        kinds = [jl_kind(t) for t in split_top(inner) if t]
        nvalues = None if any(a.endswith("...") for a in args[3:]) else len(args) - 3
        calls.append((m.group(1), kinds, nvalues))
    return calls


def test_every_ccall_matches_its_prototype():
    protos = header_prototypes()
    calls = shim_ccalls()
    assert len(calls) >= 25
    for name, kinds, nvalues in calls:
        assert name in protos, "%s is not declared in include/*.h" % name
        assert kinds == protos[name], "%s: shim passes %s, header declares %s" % (name, kinds, protos[name])
        if nvalues is not None:
            assert nvalues == len(kinds), "%s: %d values for %d declared argument types" % (name, nvalues, len(kinds))


def test_shim_covers_the_reference_surface():
    """src/SpaceCharge.jl:17 exports + the functions SURVEY.md 8(b) lists; every product entry point has a binding."""
    text = open(SHIM).read()
    for exported in ("Mesh3D", "deposit!", "clear_mesh!", "interpolate_field", "solve!"):
        assert re.search(r"^export .*\b%s" % re.escape(exported), text, flags=re.M), exported
    for fn in ("solve_freespace!", "get_green_function!", "field_green_function", "potential_green_function"):
        assert re.search(r"\b%s\(" % re.escape(fn), text), fn
    bound = {name for name, _, _ in shim_ccalls()}
    product = {n for n in header_prototypes() if not n.startswith("scb_debug_")}
    # bookkeeping calls with no use from Julia (timing/launch counters, explicit stream changes, version, cache control)
    optional = {"scb_version", "scb_enable_timing", "scb_get_timing", "scb_launch_count", "scb_workspace_bytes",
                "scb_set_stream", "scb_drop_green_cache", "scb_sync", "scb_cell_index", "scb_bounds_strided",
                "scb_step_strided"}
    missing = product - bound - optional
    assert not missing, "no Julia binding for %s" % sorted(missing)
