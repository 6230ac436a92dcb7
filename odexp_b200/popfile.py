"""Parser for odexp model files (.pop) — the model-file front end of the B200 population integrator.

Written from the language description in the reference's manual (README.md "MODEL FILE" section) and
the behaviour of its generator `src/odexp` (prescan :311-402, regexes :26-142): keywords are matched
case-insensitively by prefix (PAR, EXPR, AUX, D<var>/DT, INIT, LET, MACRO, TIMESPAN, OPT, FUN, %M, %C,
%BIRTH, %DEATH, %REPLI), `#` starts a comment, `{...}` carries attributes, LET constants are substituted
textually, `;` separates several PAR / OPT entries on one line, a bare `name value` line is a parameter.

Accepted beyond the basic keywords: `decl const` / `const` arrays, FILE tables, LAG chains, array entries with one or
more indices `name[i=a:b][j=c:d]`, multi-line `FUN ... end` blocks.  Constructs outside the population path (`decl`
of anything but constants and delay stacks, the TIMESPAN array mini-language) raise PopError with the line.
"""
import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

IDENT = r"[^\W\d][\w]*"            # unicode letters allowed, like the reference's $re_identifier


class PopError(ValueError):
    pass


@dataclass
class Entry:
    name: str
    rhs: str
    attr: str = ""
    length: int = 0          # 0 = scalar, >0 = 1-D array of that many elements
    iterator: str = ""       # iterator name of an array entry
    start: int = 0
    dims: List[Tuple[str, int, int]] = field(default_factory=list)      # (iterator, first, end) of every index of `name[i=a:b][j=c:d]`;
                                                                        # length = number of entries, iterator / start = the first index


@dataclass
class Fun:
    name: str
    args: List[str]
    body: str                # expression (one-liner) — `return` is added by the emitters; statements when block is set
    block: bool = False      # `FUN name(args)` ... `end`: C statements, printed as the function body (src/odexp:676-737)


@dataclass
class Coupling:
    name: str
    rhs: str
    attr: str = ""
    lowrank: Optional[Dict[str, str]] = None    # {var, fun, U, V, rank} for lrwexp / {var, fun} for lrexp


@dataclass
class PopModel:
    name: str
    pars: List[Entry] = field(default_factory=list)
    macros: List[Tuple[str, str]] = field(default_factory=list)
    exprs: List[Entry] = field(default_factory=list)
    auxs: List[Entry] = field(default_factory=list)
    eqs: List[Entry] = field(default_factory=list)
    inits: List[Entry] = field(default_factory=list)
    mfds: List[Entry] = field(default_factory=list)
    cpls: List[Coupling] = field(default_factory=list)
    funs: List[Fun] = field(default_factory=list)
    birth: Optional[str] = None
    death: Optional[str] = None
    repli: Optional[str] = None
    options: Dict[str, str] = field(default_factory=dict)
    timespan: List[float] = field(default_factory=list)
    consts: List[Tuple[str, str, str, str]] = field(default_factory=list)      # (C type, name, "[..]" or "", initialiser)
    stacks: List[Tuple[str, int]] = field(default_factory=list)                # `decl stack name[K]`: history stacks of delay() (name, K; K = 0: a scalar)
    base_dir: str = "."       # FILE tables are read relative to the model file

    # flattened dimensions (arrays count element-wise, as in the reference's nbr_expr etc.)
    @staticmethod
    def _count(entries):
        return sum(max(1, e.length) for e in entries)

    @property
    def nvar(self): return self._count(self.eqs)
    @property
    def nexpr(self): return self._count(self.exprs)
    @property
    def naux(self): return self._count(self.auxs)
    @property
    def npsi(self): return len(self.cpls)
    @property
    def nmfd(self): return len(self.mfds)
    @property
    def npar(self): return len(self.pars)

    def flat_names(self, entries):
        out = []
        for e in entries:
            if len(e.dims) > 1:
                import itertools
                out += [e.name + "".join("[%d]" % k for k in idx) for idx in itertools.product(*[range(a, b) for _, a, b in e.dims])]
            elif e.length:
                out += ["%s[%d]" % (e.name, e.start + k) for k in range(e.length)]
            else:
                out.append(e.name)
        return out

    def index_of(self, entries, name):
        """flattened base index of entry `name`"""
        k = 0
        for e in entries:
            if e.name == name:
                return k
            k += max(1, e.length)
        raise KeyError(name)


# option abbreviations -> long names (reference src/options.c:6-79), restricted to the ones the path reads
OPT_ABBREV = {"r": "res", "meth": "solver", "ps": "popsize", "pm": "popmode", "pr": "progress", "ld": "loudness",
              "rs": "reseed", "p": "particle"}
OPT_DEFAULTS = {"res": "201", "hmin": "1e-5", "h0": "1e-1", "abstol": "1e-6", "reltol": "1e-6", "solver": "rkck",
                "lrabstol": "1e-6", "lrreltol": "1e-6", "lrmax": "31", "popmode": "population", "popsize": "1",
                "ssahmin": "0.0", "aleap": "0.01", "seed": "3141592", "reseed": "1", "init": "0"}

_RE_ARRAY = re.compile(r"^(%s)\s*\[\s*(\w+)\s*=\s*(\d+)\s*:\s*(\d+)\s*\]$" % IDENT)
_RE_INDEX = re.compile(r"\[\s*(\w+)\s*=\s*(\d+)\s*:\s*(\d+)\s*\]")


def _array_entry(name, brackets, rhs, attr, lineno, raw):
    """`name[i=a:b]` or `name[i=a:b][j=c:d]...` (declare_scalar_or_array, src/odexp:1288-1340: `double name[b][d]`, nested loops in the
    order of the indices, entries stored row-major over the iterated ranges)"""
    dims = [(mm.group(1), int(mm.group(2)), int(mm.group(3))) for mm in _RE_INDEX.finditer(brackets)]
    if not dims or _RE_INDEX.sub("", brackets).strip():
        raise PopError("line %d: unsupported index expression in '%s'" % (lineno, raw.strip()))
    if any(b <= a for _, a, b in dims):
        raise PopError("line %d: empty iterator range in '%s'" % (lineno, raw.strip()))
    if len({it for it, _, _ in dims}) != len(dims):
        raise PopError("line %d: the indices of '%s' need distinct iterators" % (lineno, raw.strip()))
    n = 1
    for _, a, b in dims:
        n *= b - a
    return Entry(name, rhs, attr, length=n if len(dims) > 1 else dims[0][2] - dims[0][1], iterator=dims[0][0], start=dims[0][1], dims=dims)


def _strip_comment(line):
    # '#' starts a comment (model expressions never contain '#')
    i = line.find("#")
    return line if i < 0 else line[:i]


def _split_attr(s):
    m = re.search(r"\{(.*?)\}\s*$", s)
    if m:
        return s[:m.start()].rstrip(), m.group(1).strip()
    return s.rstrip(), ""


def _name_and_rhs(rest, lineno, raw):
    """`name[iter=a:b] rhs {attr}` -> Entry"""
    rest, attr = _split_attr(rest)
    m = re.match(r"^(%s(?:\s*\[[^\]]*\])*)\s+(.*)$" % IDENT, rest, flags=re.S)
    if not m:
        raise PopError("line %d: cannot parse '%s'" % (lineno, raw.strip()))
    lhs, rhs = m.group(1).replace(" ", ""), m.group(2).strip()
    if "[" in lhs:
        return _array_entry(lhs[:lhs.index("[")], lhs[lhs.index("["):], rhs, attr, lineno, raw)
    return Entry(lhs, rhs, attr)


def parse(text: str, name: str = "model", base_dir: str = ".") -> PopModel:
    m = PopModel(name=name)
    m.base_dir = base_dir
    # pass 1: join continuation lines, collect LET constants
    lines, cont = [], ""
    for raw in text.splitlines():
        s = raw.rstrip("\n")
        if re.search(r"(\\|\.\.\.)\s*$", s) and not s.lstrip().startswith("#"):
            cont += re.sub(r"(\\|\.\.\.)\s*$", " ", s)
            continue
        lines.append(cont + s)
        cont = ""
    if cont:
        lines.append(cont)
    # LAG name {root, len, mean, init}: rewritten in place into a chain of hidden dynamical variables, the auxiliary variable at
    # the end of the chain and their initial conditions, exactly like the reference's preprocessor (src/odexp:383-391)
    expanded = []
    for s in lines:
        ml = re.match(r"^\s*lag\w*\s+(%s)\s*\{(.*?)\}" % IDENT, _strip_comment(s), flags=re.I)
        if not ml:
            expanded.append(s)
            continue
        v, attr = ml.group(1), ml.group(2)
        get = lambda key: (re.search(r"\b%s\s*=\s*(.*?)([,\s]|$)" % key, attr) or [None, None])[1]
        root, ic, ln, mean = get("root"), get("init"), get("len"), get("mean")
        if None in (root, ic, ln, mean):
            raise PopError("LAG %s needs root, len, mean and init attributes" % v)
        expanded += ["d%s_chain[i=0:%s]/dt = linchaindelay(%s,%s_chain,i,%s,%s) {hidden}" % (v, ln, root, v, mean, ln),
                     "macro %s_end %s-1" % (v, ln), "AUX %s %s_chain[%s_end]" % (v, v, v), "INIT %s_chain[i=0:%s] %s {hidden}" % (v, ln, ic)]
    lines = expanded
    # multi-line functions: `FUN name(args)` without `=`, statements, `end` -> one line `fun name(args) = {block} statements`
    joined, block = [], None
    for s in lines:
        t = _strip_comment(s).strip()
        if block is not None:
            if re.match(r"^end\b", t, flags=re.I):
                joined.append(block[0] + " = \x01" + " ".join(block[1:]))
                block = None
            elif t:
                block.append(t)
            continue
        if re.match(r"^fun\w*\s+%s\s*\([^)]*\)\s*$" % IDENT, t, flags=re.I):
            block = [t]
            continue
        joined.append(s)
    if block is not None:
        raise PopError("FUN %s: multi-line function without `end`" % block[0])
    lines = joined
    lets = []
    for s in lines:
        f = _strip_comment(s).split()
        if f and f[0].lower().startswith("let"):
            if len(f) < 3:
                raise PopError("LET needs a name and a value: '%s'" % s.strip())
            lets.append((f[1], " ".join(f[2:])))
    # pass 2: substitute LETs, split PAR/OPT at ';', dispatch on keyword
    for lineno, raw in enumerate(lines, 1):
        s = _strip_comment(raw).strip()
        if not s:
            continue
        if s.lower().startswith("let"):
            continue
        for k, v in lets:
            s = re.sub(r"\b%s\b" % re.escape(k), v, s)
        if re.match(r"^%s\s+[-+]?[0-9]*\.?[0-9]+([eE][-+]?[0-9]+)?\s*(\{.*\})?$" % IDENT, s) and \
                not re.match(r"^(par|expr|aux|init|mac|time|opt|fun|lag|cmd|const|decl|file|fi)\w*\s", s, flags=re.I):
            s = "par " + s                                    # naked parameter
        kw = s.split()[0]
        kwl = kw.lower()
        rest = s[len(kw):].strip()
        if kwl.startswith("par") or kwl.startswith("opt"):
            # the reference breaks PAR / OPT lines at every ';' before it looks for comments (src/odexp:397), so a comment ends at the
            # next ';' and what follows is another entry: `PAR a 1 {int} # an integer; b 2.3` declares a and b
            full = raw.strip()
            i = full.find("#")
            if i >= 0 and ";" in full[i:] and full.lower().startswith(kwl):
                for k, v in lets:
                    full = re.sub(r"\b%s\b" % re.escape(k), v, full)
                rest = full[len(kw):]
            parts = [p for p in (_strip_comment(p).strip() for p in rest.split(";")) if p]
        else:
            parts = [rest]
        for rest in parts:
            _dispatch(m, kwl, kw, rest, lineno, raw)
    _validate(m)
    return m


def _dispatch(m, kwl, kw, rest, lineno, raw):
    if kwl.startswith("par"):
        e = _name_and_rhs(rest, lineno, raw)
        if e.length:
            raise PopError("line %d: array parameters are not supported" % lineno)
        m.pars.append(e)
    elif kwl.startswith("expr"):
        m.exprs.append(_name_and_rhs(rest, lineno, raw))
    elif kwl.startswith("aux"):
        e = _name_and_rhs(rest, lineno, raw)
        m.auxs.append(e)          # {vect}: the rhs is a statement that fills the array (`aux arr[i=0:5] myfunc(arr) {vect}`)          # scalar, or one-dimensional array `aux a[i=0:5] X[i]*X[i]` (define_array_functions, src/odexp:1381-1436)
    elif kwl.startswith("init"):
        m.inits.append(_name_and_rhs(rest, lineno, raw))
    elif kwl.startswith("mac"):
        f = rest.split(None, 1)
        if len(f) != 2:
            raise PopError("line %d: MACRO needs a name and a value" % lineno)
        m.macros.append((f[0], f[1].strip()))
    elif kwl.startswith("time"):
        try:
            m.timespan = [float(x) for x in rest.split()]
        except ValueError:
            raise PopError("line %d: TIMESPAN must be numeric here (array mini-language is host-only)" % lineno)
    elif kwl.startswith("opt"):
        f = rest.split(None, 1)
        if len(f) != 2:
            raise PopError("line %d: OPT needs a name and a value" % lineno)
        m.options[OPT_ABBREV.get(f[0], f[0])] = f[1].strip()
    elif kwl.startswith("fun"):
        mm = re.match(r"^(%s)\s*\(([^)]*)\)\s*=\s*(.*?);?\s*$" % IDENT, rest, flags=re.S)
        if not mm:
            raise PopError("line %d: cannot parse the function definition '%s'" % (lineno, raw.strip()))
        body = mm.group(3).strip()
        if body.startswith("\x01"):
            m.funs.append(Fun(mm.group(1), [a.strip() for a in mm.group(2).split(",") if a.strip()], body[1:].strip() + ";", block=True))
            return
        # the reference's prescan (src/odexp:356-363) rewrites FUN lines before its generator and its loader read them back from
        # .odexp/equations.pop: `fun f(a) = return expr;` for a one-liner, `fun f(a) = stmt;stmt;return x;` (the lines of the block glued
        # together, `end` dropped) for a multi-line function.  Both spellings are accepted, so the device model can be built from that file.
        one = re.match(r"^return\b\s*(.*)$", body, flags=re.S)
        if one and ";" not in one.group(1) and not re.search(r"\b(for|while|if|return)\b", one.group(1)):
            body = one.group(1).strip()
        elif re.search(r"\b(for|while|if|return)\b", body) or ";" in body:
            glued = re.sub(r"([;{}])(?=\S)", r"\1 ", body)          # the spacing of the multi-line spelling: one statement per line, joined by blanks
            glued = re.sub(r";\s+\)", ";)", glued)
            m.funs.append(Fun(mm.group(1), [a.strip() for a in mm.group(2).split(",") if a.strip()], glued.strip() + ";", block=True))
            return
        m.funs.append(Fun(mm.group(1), [a.strip() for a in mm.group(2).split(",") if a.strip()], body))
    elif kwl == "%m" or kwl.startswith("%m"):
        e = _name_and_rhs(rest, lineno, raw)
        m.mfds.append(e)          # {vect} / {noloop}: evaluated once instead of averaged over the particles (Entry.attr)
    elif kwl.startswith("%c"):
        rest2, attr = _split_attr(rest)
        f = rest2.split(None, 1)
        lr = None
        if re.search(r"\blrw?exp\b", rest2) or re.search(r"\blrw?exp\b", attr):
            d = dict((k.strip(), v.strip()) for k, v in (kv.split("=") for kv in attr.split(",") if "=" in kv))
            d["kind"] = "lrwexp" if re.search(r"\blrwexp\b", rest2 + " " + attr) else "lrexp"
            lr = d
            m.cpls.append(Coupling(f[0], "", attr, lr))
        else:
            if len(f) != 2:
                raise PopError("line %d: %%C needs a name and an expression" % lineno)
            m.cpls.append(Coupling(f[0], f[1].strip(), attr, None))
    elif kwl.startswith("%birth"):
        m.birth = _split_attr(rest)[0]
    elif kwl.startswith("%death"):
        m.death = _split_attr(rest)[0]
    elif kwl.startswith("%repli"):
        m.repli = _split_attr(rest)[0]
    elif (kwl.startswith("dec") and "/" not in kw) or kwl.startswith("const"):      # `decay/dt = ...` is an equation, not a declaration
        # `decl const double a[] = {..}` is printed verbatim at file scope by the reference (add_declarations, src/odexp:660-674);
        # `const a[] {..}` is the older spelling the shipped examples/Nbody.pop still uses (declare_constant_array, src/odexp:643-658).
        # Only constant scalars / arrays are taken: they become constexpr locals of every generated function.
        ms = re.match(r"^stack\s+(%s)\s*(?:\[\s*([^\]]+?)\s*\])?\s*;?\s*$" % IDENT, rest) if kwl.startswith("dec") else None
        if ms:
            # `decl stack hist[NP]`: printed verbatim at file scope by the reference; the history of one delay() call site per particle
            # (src/odexp.h:151-162, src/utils_odexp.c:205-240)
            try:
                cnt = int(eval(ms.group(2), {"__builtins__": {}}, {})) if ms.group(2) else 0
            except Exception:
                raise PopError("line %d: the size of a `decl stack` array must be an integer constant" % lineno)
            m.stacks.append((ms.group(1), cnt))
            return
        mm = re.match(r"^(?:static\s+)?const\s+(double|int|long)\s+(%s)\s*((?:\[[^\]]*\])*)\s*=\s*(.+?);?\s*$" % IDENT, rest, flags=re.S) \
            if kwl.startswith("dec") else re.match(r"^()(%s)\s*((?:\[[^\]]*\])*)\s*=?\s*(.+?);?\s*$" % IDENT, rest, flags=re.S)
        if not mm:
            raise PopError("line %d: only `decl const double|int|long name[..] = value` and `decl stack name[K]` declarations are supported" % lineno)
        if bool(mm.group(3)) != mm.group(4).strip().startswith("{"):
            raise PopError("line %d: a constant array needs a braced initialiser list, a constant scalar a single value" % lineno)
        m.consts.append((mm.group(1) or "double", mm.group(2), mm.group(3).replace(" ", ""), mm.group(4).strip()))
    elif kwl.startswith("fi"):
        # FILE name filename {nrow, ncol, sep, skip}: a constant table read at generation time (declare_array_from_file,
        # src/odexp:581-641) -> `static const double name[rows][cols]`, here one more constexpr constant
        import os
        body, attr = _split_attr(rest)
        f = body.split(None, 1)
        if len(f) != 2:
            raise PopError("line %d: FILE needs a name and a file name" % lineno)
        fname = f[1].strip().strip('"')
        def opt(key, default):
            mo = re.search(r"\b%s\b\s*=\s*(\"(.*?)\"|\d+)" % key, attr, flags=re.I)
            return default if not mo else (mo.group(2) if mo.group(2) is not None else int(mo.group(1)))
        sep, skip = opt("sep", " "), opt("skip", 0)
        path = fname if os.path.isabs(fname) else os.path.join(m.base_dir, fname)
        if not os.path.exists(path) and os.path.exists(fname):
            path = fname              # the reference's generator opens the file relative to the working directory (src/odexp:581-641): the
                                      # case of a model built from its prescanned copy .odexp/equations.pop
        try:
            lines_ = open(path).read().splitlines()[skip:]
        except OSError as ex:
            raise PopError("line %d: FILE %s: %s" % (lineno, fname, ex))
        rows = [ln.split(sep) for ln in lines_ if ln.split(sep) != ['']]
        nrow, ncol = min(opt("nrow", len(rows)), len(rows)), min(opt("ncol", 10 ** 9), max(len(r) for r in rows) if rows else 0)
        if nrow < 1 or ncol < 1:
            raise PopError("line %d: FILE %s holds no data" % (lineno, fname))
        init = "{" + ", ".join("{" + ",".join(x.strip() for x in r[:ncol]) + "}" for r in rows[:nrow]) + "}"
        m.consts.append(("double", f[0], "[%d][%d]" % (nrow, ncol), init))
    elif kwl.startswith("cmd") or kwl.startswith("end"):
        return
    elif re.match(r"^d%s\s*\[" % IDENT, kw + " " + rest, flags=re.I) and re.match(r"^d(%s)\s*((?:\[[^\]]*\]\s*)+)/dt\s*=?\s*(.*)$" % IDENT, kw + " " + rest, flags=re.I | re.S):
        # array of dynamical variables `dX[i=0:5]/dt = rhs`: one equation per index, the iterator is visible in the rhs (src/odexp:1439-1506)
        mm = re.match(r"^d(%s)\s*((?:\[[^\]]*\]\s*)+)/dt\s*=?\s*(.*)$" % IDENT, kw + " " + rest, flags=re.I | re.S)
        rhs, attr = _split_attr(mm.group(3))
        e = _array_entry(mm.group(1), mm.group(2).replace(" ", ""), rhs.strip(), attr, lineno, raw)
        m.eqs.append(e)
        mi = re.search(r"\binit\s*=\s*([^,}]+)", attr)               # {init = value}: `INIT name[i=a:b] value` (src/odexp:377-382)
        if mi:
            m.inits.append(Entry(e.name, mi.group(1).strip(), "from {init = ..}", length=e.length, iterator=e.iterator, start=e.start, dims=list(e.dims)))
    elif re.match(r"^d(%s)/dt$" % IDENT, kw, flags=re.I):
        var = re.match(r"^d(%s)/dt$" % IDENT, kw, flags=re.I).group(1)
        rhs, attr = _split_attr(rest)
        rhs = rhs.lstrip("=").strip()
        m.eqs.append(Entry(var, rhs, attr))
        mi = re.search(r"\binit\s*=\s*([^,}]+)", attr)
        if mi:
            m.inits.append(Entry(var, mi.group(1).strip(), "from {init = ..}"))
    else:
        mm = re.match(r"^d(%s)/dt\s*=?\s*(.*)$" % IDENT, (kw + " " + rest), flags=re.I | re.S)
        if mm:                                                   # `dx/dt=-x`: no blank before the `=`
            rhs, attr = _split_attr(mm.group(2))
            m.eqs.append(Entry(mm.group(1), rhs.strip(), attr))
            mi = re.search(r"\binit\s*=\s*([^,}]+)", attr)
            if mi:
                m.inits.append(Entry(mm.group(1), mi.group(1).strip(), "from {init = ..}"))
            return
        raise PopError("line %d: unknown or unsupported keyword '%s' in '%s'" % (lineno, kw, raw.strip()))


def _validate(m):
    if not m.eqs:
        raise PopError("model has no dynamical variable (d<name>/dt = ...)")
    # the prescanned copy of a model file (.odexp/equations.pop) carries both the `{init = ..}` attribute and the INIT line the reference's
    # prescan made of it (src/odexp:377-382): the explicit line stands
    explicit = {e.name for e in m.inits if e.attr != "from {init = ..}"}
    m.inits = [e for e in m.inits if not (e.attr == "from {init = ..}" and e.name in explicit)]
    dyn = [e.name for e in m.eqs]
    ini = [e.name for e in m.inits]
    if dyn != ini:
        raise PopError("initial conditions must be declared for, and in the order of, the dynamical variables: %s vs %s" % (dyn, ini))
    for e, i in zip(m.eqs, m.inits):
        if (e.length, e.start, [d[1:] for d in e.dims]) != (i.length, i.start, [d[1:] for d in i.dims]):
            raise PopError("initial condition of %s does not cover the same index range as its equation" % e.name)
    for k, v in OPT_DEFAULTS.items():
        m.options.setdefault(k, v)
    if not m.timespan:
        m.timespan = [0.0, 1.0]


def parse_file(path: str) -> PopModel:
    import os
    with open(path, encoding="utf8") as f:
        return parse(f.read(), name=os.path.splitext(os.path.basename(path))[0], base_dir=os.path.dirname(os.path.abspath(path)))
