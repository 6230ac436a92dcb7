"""CUDA emitter: PopModel -> model.cu (device + host functions of one population model).

Counterpart of the reference generator's C emitter (`src/odexp`: print_ode_rhs :981-1043, set_population_meanfields
:874-903, set_population_coupling :905-954, set_population_birth_rate :860-872, print_ode_init_cond :1045-1069,
print_header_defines :1071-1195).  Instead of a CPU `model.c` whose macros search variable names at run time
(src/odexp.h:218-364), it emits `struct OdxModel` with compile-time dimensions and fully inlinable
`__host__ __device__` functions; `odx_engine.cuh` instantiates the stage kernels around it.

Scope rules follow the reference:
  * inside D/DT, AUX, %DEATH, %REPLI: parameters (unless {impl}/{init}), EXPR, %C, %M, AUX and the dynamical variables
    are local names; everything else is reached through MU()/MY()/ME()/MA()/MF()/MC().
  * inside %M, %C, %BIRTH: only the macros exist (MY/ME/MA/MU/MF, OY/OE for the other particle).
  * inside EXPR / INIT (initial conditions): parameters tagged {init}/{ever}, previously defined EXPR, SY/SE (sister).
"""
import re
from .popfile import PopModel, PopError, IDENT

_C_KEYWORDS = {"double", "int", "float", "long", "return", "if", "else", "for", "while", "const", "static", "void", "char"}


# identifiers a C model may use freely (the reference compiles model.c as C) that the generated C++ / CUDA cannot: the template
# parameter of the generated functions and the C++-only keywords
_CXX_RESERVED = {"C", "alignas", "alignof", "and", "and_eq", "bitand", "bitor", "catch", "class", "compl", "const_cast", "constexpr",
                 "decltype", "delete", "dynamic_cast", "explicit", "export", "friend", "mutable", "namespace", "new", "noexcept", "not", "not_eq",
                 "nullptr", "operator", "or", "or_eq", "private", "protected", "public", "reinterpret_cast", "static_assert", "static_cast",
                 "template", "this", "thread_local", "throw", "try", "typeid", "typename", "using", "virtual", "wchar_t", "xor", "xor_eq"}
# functions of the CUDA math API that live in the global namespace and are not part of C's libm: a model FUN of that name would redefine them
_CUDA_MATH = {"max", "min", "abs", "rsqrt", "rcbrt", "sinpi", "cospi", "sincospi", "sincos", "exp10", "erfinv", "erfcinv", "erfcx", "normcdf",
              "normcdfinv", "cyl_bessel_i0", "cyl_bessel_i1", "norm3d", "norm4d", "rnorm3d", "rnorm4d", "rhypot", "norm", "rnorm", "saturate",
              "umin", "umax", "llmin", "llmax", "ullmin", "ullmax", "fdivide", "signbit", "clock", "clock64",
              # libm: math.h declares them extern "C"; a model FUN of the same name defines (and in C overrides) the function, in C++ the
              # definition clashes with the declaration
              "gamma", "lgamma", "tgamma", "j0", "j1", "jn", "y0", "y1", "yn", "erf", "erfc", "expm1", "log1p", "logb", "ilogb", "cbrt", "hypot",
              "fdim", "fmax", "fmin", "fma", "fmod", "round", "lround", "trunc", "rint", "lrint", "nearbyint", "remainder", "remquo", "copysign",
              "nan", "exp2", "log2", "log10", "scalbn", "scalbln", "ldexp", "frexp", "modf", "significand", "drem", "finite", "isinf", "isnan",
              "sin", "cos", "tan", "asin", "acos", "atan", "atan2", "sinh", "cosh", "tanh", "asinh", "acosh", "atanh", "exp", "log", "pow",
              "sqrt", "ceil", "floor", "fabs", "nextafter", "div", "labs", "rand", "srand", "exit", "abort", "index", "time", "select", "link",
              "signal", "pause", "sleep", "read", "write", "open", "close", "pipe", "dup", "sync", "nice", "access", "alarm", "truncate"}
_fun_renames = set()          # the FUN names of the model being emitted that collide with _CUDA_MATH / libc (set by Emitter.__init__)


def mangle(name: str) -> str:
    """ASCII-safe identifier (nvcc device code does not take the UTF-8 identifiers gcc accepts) that is neither a C++ keyword, the
    generated functions' template parameter, nor a model function named like a CUDA builtin"""
    out = []
    for ch in name:
        out.append(ch if (ch.isascii() and (ch.isalnum() or ch == "_")) else "_u%04x" % ord(ch))
    s = "".join(out)
    if s in _CXX_RESERVED or s in _fun_renames:
        s = "odx_v_" + s
    return s


def _mangle_expr(expr: str) -> str:
    def rep(m):
        name = m.group(0)
        if name in ("max", "min") and name not in _fun_renames and re.match(r"\s*\(", expr[m.end():]):
            return "odx_" + name          # the reference's type-generic max(a,b) / min(a,b) macros (src/odexp.h:204-212), see _MINMAX
        if name == "abs" and name not in _fun_renames and re.match(r"\s*\(", expr[m.end():]):
            return "odx_iabs"             # C's abs() is int abs(int): abs(-0.5) is 0 in the reference's model.c; C++ would pick fabs
        return mangle(name)
    return re.sub(IDENT, rep, expr)


# max(a,b) / min(a,b) of src/odexp.h:204-212 (`_a > _b ? _a : _b` on the arguments' own types: the usual arithmetic conversions, a NaN in
# the first argument gives the second, unlike fmax / CUDA's max; `+ 0` makes the return type a value, not a reference to an argument); printed into the generated file only when the model calls them
_MINMAX = """template <class OdxA, class OdxB> ODX_HD auto odx_max(OdxA a, OdxB b) -> decltype((a > b ? a : b) + 0) { return a > b ? a : b; }
template <class OdxA, class OdxB> ODX_HD auto odx_min(OdxA a, OdxB b) -> decltype((a < b ? a : b) + 0) { return a < b ? a : b; }
"""


def _find_call(expr, fname, start=0):
    """yield (begin, end, [args]) for calls fname(...) with balanced parentheses"""
    pat = re.compile(r"\b%s\s*\(" % re.escape(fname))
    pos = start
    while True:
        m = pat.search(expr, pos)
        if not m:
            return
        i = m.end()
        depth, args, cur = 1, [], []
        while i < len(expr) and depth:
            ch = expr[i]
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
                if depth == 0:
                    break
            if ch == "," and depth == 1:
                args.append("".join(cur)); cur = []
            else:
                cur.append(ch)
            i += 1
        if depth:
            raise PopError("unbalanced parentheses in '%s'" % expr)
        args.append("".join(cur))
        yield m.start(), i + 1, [a.strip() for a in args]
        pos = i + 1


def _pow_peephole(expr: str) -> str:
    """pow(x, c) with a literal c in {0.5, 1.5, 2, 3} -> multiplications / sqrt (FP64 pow costs ~100 instructions)"""
    table = {"0.5": "odx_pow05", "1.5": "odx_pow15", "2": "odx_pow2", "2.0": "odx_pow2", "3": "odx_pow3", "3.0": "odx_pow3"}
    changed = True
    while changed:
        changed = False
        for b, e, args in _find_call(expr, "pow"):
            if len(args) == 2 and args[1] in table:
                expr = expr[:b] + "%s(%s)" % (table[args[1]], args[0]) + expr[e:]
                changed = True
                break
    return expr


def _strip_outer_parens(e):
    e = e.strip()
    while e.startswith("(") and e.endswith(")"):
        depth = 0
        for i, ch in enumerate(e):
            depth += ch == "("
            depth -= ch == ")"
            if depth == 0 and i < len(e) - 1:
                return e
        e = e[1:-1].strip()
    return e


def _ternary_to_select(expr: str, double_funs=None) -> str:
    """`c ? a : b` (whole expression) -> odx_sel(c, a, b): both arms are evaluated, the result is a select.
    On the GPU this removes a divergent branch per use and lets the compiler share sub-expressions between several
    guarded terms (the N-body fixture guards each force component separately around the same r^3).  Only applied when
    neither arm draws random numbers or raises STOPIF (evaluating those eagerly would change behaviour)."""
    e = _strip_outer_parens(expr)
    depth, q = 0, -1
    for i, ch in enumerate(e):
        depth += ch in "(["
        depth -= ch in ")]"
        if ch == "?" and depth == 0:
            q = i
            break
    if q < 0:
        return expr
    depth, nest, c = 0, 0, -1
    for i in range(q + 1, len(e)):
        ch = e[i]
        depth += ch in "(["
        depth -= ch in ")]"
        if depth == 0 and ch == "?":
            nest += 1
        elif depth == 0 and ch == ":":
            if nest == 0:
                c = i
                break
            nest -= 1
    if c < 0:
        return expr
    cond, a, b = e[:q].strip(), e[q + 1:c].strip(), e[c + 1:].strip()
    if re.search(r"\b(rand01|randN|exprand|poisson|DWDT|STOPIF%s)\b" % "".join("|" + re.escape(n) for n in getattr(double_funs, "rng_names", ())), a + " " + b):
        return expr
    if re.search(r"\[[^\]]*[A-Za-z_][^\]]*\]", a + " " + b):
        return expr            # an arm indexes an array with a computed index (`(i == 0) ? a : X[i-1]`): the guard may be what keeps it in range
    if double_funs is not None and not (_is_double_valued("(" + a + ")", double_funs) or _is_double_valued("(" + b + ")", double_funs)):
        return expr            # `(ID > 2) ? 3 : 1` is an int in C (its quotient by 2 an integer division); odx_sel would make it a double
    return "odx_sel(%s, %s, %s)" % (cond, _ternary_to_select(a, double_funs), _ternary_to_select(b, double_funs))


def _ternary_to_select_nested(expr: str, double_funs=None) -> str:
    """apply _ternary_to_select to every parenthesised sub-expression that is a whole ternary (after function inlining the
    guards of the model functions sit inside parentheses)"""
    out, i, brackets = [], 0, 0
    while i < len(expr):
        if expr[i] == "(":
            depth, j = 1, i + 1
            while j < len(expr) and depth:
                depth += expr[j] == "("
                depth -= expr[j] == ")"
                j += 1
            inner = _ternary_to_select_nested(expr[i + 1:j - 1], double_funs)
            conv = inner if brackets else _ternary_to_select(inner, double_funs)      # inside an array subscript the value must stay an integer
            out.append("(" + conv + ")")
            i = j
        else:
            brackets += expr[i] == "["
            brackets -= expr[i] == "]"
            out.append(expr[i])
            i += 1
    return "".join(out)


_DOUBLE_ATOM = re.compile(r"(?<![\w.])(\d+\.\d*|\.\d+|\d+[eE][-+]?\d+)|odx_o?(y|e|aux)\[|odx_(mfd|psi)\[|odx_c\.mu\[")


def _is_double_valued(den: str, double_funs) -> bool:
    """can the C type of this denominator (a call `f(...)` or a parenthesised expression) be shown to be double from its text?  A call
    of a function known to return a double; or a parenthesised expression without comparison / logical / conditional / modulo operators
    whose top level (plain parentheses, arguments of double-valued calls excluded from nothing, arguments of other calls ignored)
    holds a floating-point literal, a double-valued call or a state / expression / mean-field / parameter reference."""
    d = _strip_outer_parens(den.strip())
    mm = re.match(r"^(%s)\s*\(" % IDENT, d)
    if mm:
        calls = list(_find_call(d, mm.group(1)))
        if calls and calls[0][0] == 0 and calls[0][1] == len(d):
            return mm.group(1) in _DOUBLE_FUNS or mm.group(1) in double_funs
    if re.search(r"[<>=!&|?:%]", d):
        return False
    # blank out the argument lists of calls that are not known to return a double: what they read says nothing about their type
    flat, i = [], 0
    for mc in re.finditer(r"(%s)\s*\(" % IDENT, d):
        if mc.start() < i:
            continue
        name = mc.group(1)
        j, depth = mc.end(), 1
        while j < len(d) and depth:
            depth += d[j] == "("
            depth -= d[j] == ")"
            j += 1
        if name in _DOUBLE_FUNS or name in double_funs:
            return True
        flat.append(d[i:mc.start()])
        i = j
    flat.append(d[i:])
    rest = "".join(flat)
    if _DOUBLE_ATOM.search(rest):
        return True
    names = getattr(double_funs, "double_names", ())
    return any(tok in names for tok in re.findall(IDENT, re.sub(r"(?<![\w.])\d+[eE][-+]?\d+", "", rest)))


_DOUBLE_FUNS = {"pow", "sqrt", "exp", "log", "fabs", "cosh", "sinh", "tanh", "hypot", "cbrt", "odx_pow05", "odx_pow15", "odx_pow2", "odx_pow3"}


def _div_peephole(expr: str, double_funs=frozenset()) -> str:
    """`x / f(args)` -> `x * odx_rcp(f(args))` for divisions by a call of a function that returns a double (libm, the pow peepholes,
    one-line model functions).  A double division costs ~20 instructions on the GPU; as a reciprocal it becomes a pure call the
    compiler can share between terms that divide by the same quantity (the three force components of the N-body fixture divide by
    one r^3).  Result differs from the IEEE quotient by at most 1.5 ulp.  A division by a parenthesised expression or by anything else
    is left alone: `1/POP_SIZE`, `ID/(2)`, `7/(ID+1)` are integer divisions in the reference's C (POP_SIZE and ID are ints), and a
    reciprocal would silently turn them into floating-point ones."""
    out, pos = [], 0
    pat = re.compile(r"/\s*(%s)?\s*\(" % IDENT)
    start = 0
    while True:
        m = pat.search(expr, max(pos, start))
        if not m:
            out.append(expr[pos:])
            return "".join(out)
        i, depth = m.end(), 1
        while i < len(expr) and depth:
            depth += expr[i] == "("
            depth -= expr[i] == ")"
            i += 1
        if depth:
            out.append(expr[pos:])
            return "".join(out)
        if not _is_double_valued((m.group(1) or "") + "(" + expr[m.end():i - 1] + ")", double_funs):
            start = m.end()                 # leave this division alone, look at the ones behind (and inside) it
            continue
        out.append(expr[pos:m.start()] + "* odx_rcp(" + (m.group(1) or "") + "(" + _div_peephole(expr[m.end():i - 1], double_funs) + "))")
        pos = i


def _inline_funs(expr: str, funs, rng=()) -> str:
    """substitute one-line model functions textually (arguments parenthesised), innermost definitions last, so that the
    peepholes below see through `grav(...)` / `normCube(...)`; the functions are still emitted for host use"""
    for _ in range(8):                      # bounded nesting depth
        changed = False
        for f in funs:
            if f.block or any(not re.match(r"^%s$" % IDENT, a) for a in f.args):
                continue                    # statements / pointer or typed arguments: called, not inlined
            for b, e, args in _find_call(expr, f.name):
                if len(args) != len(f.args):
                    continue
                if any(re.search(r"\b(rand01|randN|exprand|poisson|DWDT|STOPIF|delay%s)\b" % "".join("|" + re.escape(n) for n in rng), a) for a in args):
                    continue                # call by value evaluates a draw once; substituted text would draw once per use of the parameter
                body = f.body
                # two-phase substitution so that an argument text cannot be captured by a later parameter name
                for k, a in enumerate(f.args):
                    body = re.sub(r"\b%s\b" % re.escape(a), "\x00%d\x00" % k, body)
                for k, a in enumerate(args):
                    body = body.replace("\x00%d\x00" % k, "(" + a + ")")
                expr = expr[:b] + "(" + body + ")" + expr[e:]
                changed = True
                break
            if changed:
                break
        if not changed:
            return expr
    return expr


def _rpow_peephole(expr: str) -> str:
    """odx_rcp(odx_pow15(x)) -> odx_rpow15(x) = rsqrt(x)^3 (one MUFU.RSQ64H + Newton instead of an IEEE sqrt and an IEEE
    reciprocal; <= 3 ulp), likewise odx_rcp(odx_pow05(x)) -> odx_rsqrt(x)"""
    for inner, repl in (("odx_pow15", "odx_rpow15"), ("odx_pow05", "odx_rsqrt")):
        changed = True
        while changed:
            changed = False
            for b, e, args in _find_call(expr, "odx_rcp"):
                if len(args) == 1:
                    a = _strip_outer_parens(args[0])
                    m = re.match(r"^%s\s*\(" % inner, a)
                    if m:
                        calls = list(_find_call(a, inner))
                        if calls and calls[0][0] == 0 and calls[0][1] == len(a):
                            expr = expr[:b] + "%s(%s)" % (repl, calls[0][2][0]) + expr[e:]
                            changed = True
                            break
    return expr


def _split_top(e: str, ops: str):
    """split e at the depth-0 occurrences of the single-character binary operators in `ops`; returns (parts, seps).
    A '+' / '-' counts as binary only when it follows an operand (not at the start, after another operator or inside a
    floating-point exponent)."""
    parts, seps, cur, depth = [], [], [], 0
    prev = ""                                   # previous non-space character at any depth
    for i, ch in enumerate(e):
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if depth == 0 and ch in ops:
            binary = True
            if ch in "+-":
                if prev == "" or prev in "*/(,+-?:<>=!&|":
                    binary = False
                elif prev in "eE" and re.search(r"(\d\.?|\.\d+)[eE]$", "".join(cur).rstrip()) and not re.search(r"[A-Za-z_]\w*$", "".join(cur).rstrip()[:-1] or " "):
                    binary = False
            if binary:
                parts.append("".join(cur)); seps.append(ch); cur = []
                prev = ch
                continue
        cur.append(ch)
        if not ch.isspace():
            prev = ch
    parts.append("".join(cur))
    return parts, seps


def _is_pure_product(e: str):
    """factors of e when e is a depth-0 product `[-]f1 * f2 * ...` (no depth-0 '+', binary '-', '/', '?', comparison), else None"""
    e = _strip_outer_parens(e)
    depth = 0
    for i, ch in enumerate(e):
        depth += ch in "(["
        depth -= ch in ")]"
        if depth == 0 and ch in "?:<>=!&|/%":
            return None
    parts, seps = _split_top(e, "+-")
    if len(parts) != 1:
        return None
    factors, _ = _split_top(e, "*")
    factors = [f.strip() for f in factors]
    if any(f == "" for f in factors):
        return None
    return factors


_J_DEPENDENT = re.compile(r"odx_oy|odx_oe|odx_oaux|odx_c\.oid|odx_c\.(rand01|randN|exprand|poisson|dwdt)|odx_stopif|\b(sin|cos)\s*\(")


def _hoist_invariant_factors(e: str):
    """%C term `a(i) * b(i,j)`: the factors that do not read the other particle leave the O(N^2) sum,
    psi_k(i) = a(i) * sum_j b(i,j) (one multiplication per particle instead of one per pair; the result differs from the
    reference's term-by-term product by the rounding of the distributive law).  Returns (scale or None, variant)."""
    factors = _is_pure_product(e)
    if not factors or len(factors) < 2:
        return None, e
    inv = [f for f in factors if not _J_DEPENDENT.search(f)]
    var = [f for f in factors if _J_DEPENDENT.search(f)]
    if not inv or not var:
        return None, e
    return " * ".join("(%s)" % f for f in inv), " * ".join(var)


def _norm(x: str) -> str:
    x = re.sub(r"\s+", "", x)
    while True:
        y = _strip_outer_parens(x)
        y = re.sub(r"\(\((odx_\w+\[\d+\])\)\)", r"(\1)", y)
        y = re.sub(r"\((odx_\w+\[\d+\])\)", r"\1", y)
        if y == x:
            return x
        x = y


def _zmul_peephole(expr: str) -> str:
    """odx_sel(A == B, 0.0, (A - B) * R) -> odx_zmul<ODX_EXACT>(odx_c, A, B, R).
    For a finite R the guarded product equals (A - B) * R (it is +-0 when A == B), so the FP64 compare and the select leave
    the hot path; R = inf / NaN (the self term of a 1/r^n kernel) is replaced by 0 on the integer pipe, and the rare case of a
    non-finite R with A != B raises the context's slow flag: the engine redoes that tile with the exact form."""
    pos = 0
    while True:
        calls = [c for c in _find_call(expr, "odx_sel", pos)]
        if not calls:
            return expr
        b, en, args = calls[0]
        done = False
        if len(args) == 3 and args[1].strip() in ("0.0", "0", "0."):
            cond = _strip_outer_parens(args[0])
            cparts = re.split(r"==", cond)
            factors = _is_pure_product(args[2])
            if len(cparts) == 2 and factors and len(factors) >= 2 and not re.search(r"[<>!&|?]", cond):
                A, B = _norm(cparts[0]), _norm(cparts[1])
                for k, f in enumerate(factors):
                    sub, seps = _split_top(_strip_outer_parens(f), "+-")
                    if len(sub) == 2 and seps == ["-"] and _norm(sub[0]) == A and _norm(sub[1]) == B:
                        rest = factors[:k] + factors[k + 1:]
                        R = " * ".join(rest)
                        if len(rest) == 1:
                            R = re.sub(r"^odx_rpow15\s*\(", "odx_rpow15_fast<ODX_EXACT>(", rest[0].strip())
                        rep = "odx_zmul<ODX_EXACT>(odx_c, %s, %s, %s)" % (cparts[0].strip(), cparts[1].strip(), R)
                        expr = expr[:b] + rep + expr[en:]
                        pos = b + len(rep)
                        done = True
                        break
        if not done:
            pos = b + len("odx_sel(")


def _antisymmetric_pair_term(term: str):
    """A %C term that changes sign, bit for bit, when the two particles are exchanged: the whole term is
    odx_zmul<..>(odx_c, odx_y[k], odx_oy[k], R) and R reads the particles only through products of two differences of the same
    dynamical variable ((y_u - oy_u) * (y_u - oy_u): (-d) * (-d) == d * d exactly), parameters, constants and t.  Returns
    (k, R, set of variable indices read) or None.  Such couplings take the symmetric all-pairs kernel (k_pair_sym): every
    unordered pair is evaluated once and added to both particles."""
    calls = list(_find_call(term, "odx_zmul<ODX_EXACT>"))
    if len(calls) != 1:
        return None
    b, en, args = calls[0]
    if term[:b].strip() != "" or term[en + 1:].strip() != "" or len(args) != 4:
        return None
    A, B, R = _norm(args[1]), _norm(args[2]), args[3]
    ma, mb = re.fullmatch(r"odx_y\[(\d+)\]", A), re.fullmatch(r"odx_oy\[(\d+)\]", B)
    if not ma or not mb or ma.group(1) != mb.group(1):
        return None
    r = re.sub(r"\s+", "", R)
    while True:
        r2 = re.sub(r"\((odx_o?y\[\d+\])\)", r"\1", r)
        r2 = re.sub(r"\(odx_y\[(\d+)\]-odx_oy\[\1\]\)", r"@D\1@", r2)
        r2 = re.sub(r"\(odx_oy\[(\d+)\]-odx_y\[\1\]\)", r"@D\1@", r2)
        r2 = re.sub(r"\((@D\d+@)\)", r"\1", r2)
        if r2 == r:
            break
        r = r2
    used = {int(ma.group(1))} | {int(x) for x in re.findall(r"@D(\d+)@", r)}
    r = re.sub(r"@D(\d+)@\*@D\1@", "@SQ@", r)
    if "@D" in r or re.search(r"odx_o?y\b|odx_o?e\b|odx_o?aux\b|odx_psi|odx_mfd|odx_stopif", r):
        return None
    if re.search(r"odx_c\.(?!mu\[|muv\[)", r):
        return None
    return int(ma.group(1)), R, used


def _nd(e):
    """an array entry with more than one index (`name[i=a:b][j=c:d]`)"""
    return len(getattr(e, "dims", ())) > 1


def _arr_decl(e):
    return mangle(e.name) + "".join("[%d]" % b for _, a, b in e.dims)


def _arr_loops(e):
    return " ".join("for (size_t %s = %d; %s < %d; ++%s)" % (mangle(it), a, mangle(it), b, mangle(it)) for it, a, b in e.dims)


def _arr_ref(e):
    return mangle(e.name) + "".join("[%s]" % mangle(it) for it, _, _ in e.dims)


def _arr_flat(e):
    """position of entry [i][j].. among the stored entries: row-major over the iterated ranges, the order of the reference's nested loops
    (`*(aux_++) = a[i][j];`, src/odexp:1381-1436)"""
    s = ""
    for k, (it, a, b) in enumerate(e.dims):
        term = "(%s - %d)" % (mangle(it), a)
        s = term if k == 0 else "(%s*%d + %s)" % (s, b - a, term)
    return s


def is_vect(e):
    """%M name rhs {vect} / {noloop}"""
    return bool(re.search(r"\b(vect|noloop)\b", getattr(e, "attr", "") or ""))


class Emitter:
    def __init__(self, m: PopModel, fast_pow: bool = True, fast_div: bool = True):
        self.m = m
        self.fast_pow = fast_pow
        self.fast_div = fast_div
        self._renames = {f.name for f in m.funs if f.name in _CUDA_MATH}
        self._activate()
        # model functions that reach the simulation through the reference's global SIM (MU(par), POP_SIZE) or draw random numbers
        # (rand01() ... are plain C functions there): here they take the evaluation context as their first argument.  rng_funs: the ones
        # that draw, directly or through another function (a call of one makes the expression a drawing one).
        self.ctx_funs, self.rng_funs = set(), set()
        for _ in range(len(m.funs) + 1):
            for f in m.funs:
                calls = set(re.findall(IDENT, f.body))
                if re.search(r"\b(rand01|randN|exprand|poisson|DWDT)\b", f.body) or calls & self.rng_funs:
                    self.rng_funs.add(f.name)
                if re.search(r"\b(MU|POP_SIZE)\b", f.body) or f.name in self.rng_funs or calls & self.ctx_funs:
                    self.ctx_funs.add(f.name)
        for f in m.funs:
            bad = re.search(r"\b(MY|ME|MA|MC|MF|OY|OE|OA|SY|SE|SA|KY|ID|MID|OID|STOPIF|ATBIRTH|ATREPLI|ISDAUGHTER|ISMOTHER|delay)\b", f.body)
            if bad:
                raise PopError("FUN %s: %s is not available inside a model function (in the reference it needs the particle in scope); "
                               "pass the value as an argument" % (f.name, bad.group(1)))
        self.draw_pat = r"\b(DWDT|rand01|randN|exprand|poisson%s)\b" % "".join("|" + re.escape(n) for n in sorted(self.rng_funs))
        self.var_idx = {e.name: m.index_of(m.eqs, e.name) for e in m.eqs}       # flattened base index (arrays count element-wise)
        self.arrays = {e.name for e in list(m.eqs) + list(m.auxs) if e.length}
        self.par_idx = {e.name: i for i, e in enumerate(m.pars)}
        self.mfd_idx = {e.name: i for i, e in enumerate(m.mfds)}
        self.psi_idx = {c.name: i for i, c in enumerate(m.cpls)}
        self.aux_idx = {e.name: m.index_of(m.auxs, e.name) for e in m.auxs}
        self.delay_sites = []
        # model functions that return a double: one-liners always do, a block only without a `return (int|long|char|float)` cast
        class _Names(frozenset):
            pass
        self.double_funs = _Names(f.name for f in m.funs if not (f.block and re.search(r"return\s*\((int|long|char|float)\)", f.body)))
        # identifiers that are doubles in the generated code: dynamical variables, expressions, auxiliaries, parameters, mean fields,
        # couplings (scalars: an array name alone is a pointer) and `decl const double` scalars; the arguments of the model functions
        dn = {e.name for e in list(m.eqs) + list(m.exprs) + list(m.auxs) if not e.length}
        dn |= {e.name for e in m.pars if not re.search(r"\b(int|long)\b", e.attr or "")}      # `par n 3 {int}` is an int in the generated code
        dn |= {e.name for e in m.mfds} | {c.name for c in m.cpls}
        dn |= {n for t, n, d, init in m.consts if "double" in t and not d}
        for f in m.funs:
            dn |= {a.strip() for a in f.args if re.fullmatch(IDENT, a.strip())}
        dn.add("t")
        self.double_funs.double_names = frozenset(dn)
        self.double_funs.rng_names = tuple(sorted(self.rng_funs))
        self.expr_base = {}
        k = 0
        for e in m.exprs:
            self.expr_base[e.name] = k
            k += max(1, e.length)

    def _activate(self):
        """mangle() is a module function: point it at this model's renamed functions (several emitters may be alive at once)"""
        if _fun_renames != self._renames:
            _fun_renames.clear()
            _fun_renames.update(self._renames)

    def _ctx_calls(self, e):
        """calls of context-taking model functions get the context as first argument"""
        for name in self.ctx_funs:
            for b, en, args in list(_find_call(e, name))[::-1]:
                inner = e[b:en]
                k = inner.index("(")
                rest = inner[k + 1:].lstrip()
                e = e[:b] + inner[:k + 1] + ("odx_c)" if rest.startswith(")") else "odx_c, " + inner[k + 1:]) + e[en:]
        return e

    def _fun_body(self, f, body):
        """the body of a context-taking model function: MU(par), POP_SIZE and the draws go through the context"""
        body = self._ctx_calls(body)
        body = re.sub(r"\bMU\(\s*(%s)\s*\)" % IDENT, lambda mm: self._macro("MU", mm.group(1), "FUN " + f.name), body)
        body = re.sub(r"\bPOP_SIZE\b", "((int)odx_c.pop_size)", body)
        body = re.sub(r"\bDWDT\b", "odx_c.dwdt()", body)
        return re.sub(r"(?<![\w.])(rand01|randN|exprand|poisson)\s*\(", r"odx_c.\1(", body)

    # ---- delay() call sites ----------------------------------------------------------------------
    def _delay_site(self, args, where):
        """register delay(t, x, tau, ic, &stack[idx]); returns (site, t expression, tau expression, stack index expression)"""
        if len(args) != 5:
            raise PopError("delay(t, x, tau, ic, &stack) takes five arguments (%s)" % where)
        mm = re.match(r"^&\s*(%s)\s*(?:\[(.+)\])?$" % IDENT, args[4].strip())
        stacks = dict(self.m.stacks)
        if not mm or mm.group(1) not in stacks:
            raise PopError("delay(): the last argument must be the address of a `decl stack` (%s)" % where)
        name, idx = mm.group(1), (mm.group(2) or "0").strip()
        if bool(mm.group(2)) != bool(stacks[name]):
            raise PopError("delay(): `%s` is %s (%s)" % (name, "an array of stacks: index it" if stacks[name] else "a single stack", where))
        if args[3] not in {f.name for f in self.m.funs}:
            raise PopError("delay(): the history function '%s' must be a FUN of one argument (%s)" % (args[3], where))
        forbidden = {x.name for x in self.m.auxs} | {x.name for x in self.m.mfds} | {x.name for x in self.m.cpls}
        for a in (args[0], args[1], idx):
            bad = set(re.findall(IDENT, a)) & forbidden
            if bad or re.search(r"\b(MA|MF|MC|OY|OE|OA|KY|rand01|randN|DWDT)\b", a):
                raise PopError("delay(): the time, the delayed value and the stack index may read t, dynamical variables, parametric expressions "
                               "and parameters only (%s reads %s): in the reference the two ode_rhs passes of one evaluation would push different "
                               "values" % (where, ", ".join(sorted(bad)) or "a macro"))
        key = (args[0], args[1], args[3], name, idx)
        for k, site in enumerate(self.delay_sites):
            if site["key"] == key:
                return k, args[0], args[2], idx
            if site["stack"] == name:
                raise PopError("delay(): the stack '%s' is used by two different call sites (%s)" % (name, where))
        self.delay_sites.append(dict(key=key, t=args[0], x=args[1], ic=args[3], stack=name, idx=idx, count=max(1, stacks[name])))
        return len(self.delay_sites) - 1, args[0], args[2], idx

    # ---- macro translation ---------------------------------------------------------------------
    def _macro(self, kind, name, where):
        name = name.strip()
        def need(d, what):
            if name not in d:
                raise PopError("%s(%s) in %s: '%s' is not a declared %s" % (kind, name, where, name, what))
            return d[name]
        if kind == "MU":
            return "(odx_c.mu[%d])" % need(self.par_idx, "parameter")
        if kind in ("MF", "MC") and (where.startswith("%C") or where.startswith("%BIRTH") or (where.startswith("%M") and "{vect}" not in where)):
            # in the reference MF() is SIM->meanfield[k] (src/odexp.h:357-364): inside the %M / %C loops of pop_rhs it reads whatever the
            # loop before it left there, a dependence on the evaluation order the population kernels do not reproduce
            raise PopError("%s(%s) in %s: mean fields / couplings cannot be read inside %%M, %%C or %%BIRTH expressions "
                           "(read them in an AUX or D/DT expression instead)" % (kind, name, where))
        if kind == "MF":
            return "(odx_mfd[%d])" % need(self.mfd_idx, "mean field")
        if kind == "MC":
            return "(odx_psi[%d])" % need(self.psi_idx, "coupling")
        if kind in ("MY", "OY", "SY", "MA", "OA", "SA") and name in self.arrays:
            raise PopError("%s(%s) in %s: macros on array variables are not supported" % (kind, name, where))
        if kind in ("MY", "OY", "SY"):
            i = need(self.var_idx, "dynamical variable")
            return {"MY": "(odx_y[%d])", "OY": "(odx_oy[%d])", "SY": "(odx_c.has_sister ? odx_c.sis_y[%d] : 0.0)"}[kind] % i
        if kind in ("ME", "OE", "SE"):
            i = need(self.expr_base, "parametric expression")
            return {"ME": "(odx_e[%d])", "OE": "(odx_oe[%d])", "SE": "(odx_c.has_sister ? odx_c.sis_e[%d] : 0.0)"}[kind] % i
        if kind == "MA" and (where.startswith("aux ") or re.match(r"^d.*/dt$", where)):
            # myself_->aux[] inside ode_rhs is what the PREVIOUS evaluation left there (the array is rewritten at the end of ode_rhs,
            # src/odexp:1030-1040): a lagged value the stage kernels do not keep
            raise PopError("MA(%s) in %s: inside AUX / D/DT expressions an auxiliary variable is read by its name (MA() there is the value "
                           "of the previous evaluation in the reference; not supported)" % (name, where))
        if kind in ("MA", "OA", "SA"):
            i = need(self.aux_idx, "auxiliary variable")
            return {"MA": "(odx_aux[%d])", "OA": "(odx_oaux[%d])", "SA": "(odx_c.has_sister ? odx_c.sis_aux[%d] : 0.0)"}[kind] % i
        raise PopError("unknown macro %s" % kind)

    def tr(self, expr: str, where: str, exact: bool = False, zmul: bool = False) -> str:
        """translate one model expression to C++.  exact = True (initial conditions): no arithmetic peepholes, so that replayed
        initial conditions stay bit-identical to the reference's"""
        self._activate()
        e = expr.strip().rstrip(";")
        bad = re.search(r"\b(FIRST|LAST|NEXT|PREVIOUS|NEXTCIRC|PREVIOUSCIRC|EVENT_TYPE|DELAY_FIRST_EVAL|SIM|myself_|other_|PRINT)\b", e)
        if bad:
            # src/odexp.h:366-412: pointers into the reference's particle list and its global simulation record
            raise PopError("%s in %s: the reference's list pointers / simulation record do not exist on the device "
                           "(neighbours: KY(name, degree); stopping: STOPIF(condition))" % (bad.group(1), where))
        if re.search(r"\brandIntMax\b", e):
            raise PopError("randIntMax() in %s: not built (src/rand_gen.c:46-63); (long)(rand01()*n) draws the same distribution" % where)
        if self.fast_div and self.m.funs and not exact:
            e = _inline_funs(e, [f for f in self.m.funs if f.name not in self.ctx_funs], rng=sorted(self.rng_funs))
        if self.ctx_funs:
            e = self._ctx_calls(e)
        e = re.sub(r"\b(MU|MF|MC|MY|OY|SY|ME|OE|SE|MA|OA|SA)\(\s*(%s)\s*\)" % IDENT,
                   lambda mm: self._macro(mm.group(1), mm.group(2), where), e)
        for b, en, args in list(_find_call(e, "KY"))[::-1]:       # KY(name, degree): the degree-th neighbour in list order, ring
            if len(args) != 2 or args[0] not in self.var_idx:
                raise PopError("KY(name, degree) in %s: '%s' is not a dynamical variable" % (where, args[0] if args else ""))
            if exact or where.startswith("%M") or where.startswith("%C") or where.startswith("%BIRTH"):
                raise PopError("KY() ring neighbours are only supported inside AUX / D/DT / %%DEATH / %%REPLI (%s)" % where)
            e = e[:b] + "odx_c.ky(%d, (%s))" % (self.var_idx[args[0]], args[1]) + e[en:]
        for b, en, args in list(_find_call(e, "delay"))[::-1]:
            # delay(t, x, tau, ic, &stack[idx]) (src/utils_odexp.c:205-240): the lookup stays in the expression, the side effects (the
            # fill of the history at t0 / while an event is pending, the push at the first evaluation of a dde step) are carried out by
            # the engine between evaluations from the call site's description (delay_sample / delay_ic below)
            if not (where.startswith("aux ") or where.startswith("d") or where == "rhs") or exact:
                raise PopError("delay() is only supported inside AUX / D/DT expressions (%s)" % where)
            e = e[:b] + "odx_c.delay_lookup(%d, (%s) - (%s), (long)(%s))" % self._delay_site(args, where) + e[en:]
        e = re.sub(r"\bDWDT\b", "odx_c.dwdt()", e)
        e = re.sub(r"\b(rand01|randN|exprand|poisson)\s*\(", lambda mm: "odx_c.%s(" % mm.group(1), e)
        e = re.sub(r"\bPOP_SIZE\b", "((int)odx_c.pop_size)", e)      # SIM->pop->size is an int in the reference (src/odexp.h): 1/POP_SIZE is 0
        e = re.sub(r"\bATBIRTH\b", "(odx_c.atbirth)", e)
        e = re.sub(r"\bATREPLI\b", "(odx_c.atrepli)", e)
        e = re.sub(r"\bISDAUGHTER\b", "(odx_c.has_sister)", e)
        e = re.sub(r"\bISMOTHER\b", "(!odx_c.has_sister)", e)
        e = re.sub(r"\bOID\b", "(odx_c.oid)", e)
        e = re.sub(r"\b(MID|ID)\b", "(odx_c.id)", e)
        for b, en, args in list(_find_call(e, "BR"))[::-1]:
            if len(args) == 2:
                e = e[:b] + "(odx_c.atbirth ? (%s) : (%s))" % (args[0], args[1]) + e[en:]
        for b, en, args in list(_find_call(e, "IST0"))[::-1]:     # IST0(at t0, afterwards), src/odexp.h:369
            if len(args) == 2:
                e = e[:b] + "(odx_c.ist0 ? (%s) : (%s))" % (args[0], args[1]) + e[en:]
        for b, en, args in list(_find_call(e, "STOPIF"))[::-1]:
            e = e[:b] + "odx_stopif(odx_c, (%s))" % args[0] + e[en:]
        if self.fast_pow and not exact:
            e = _pow_peephole(e)
        if self.fast_div and not exact:
            e = _rpow_peephole(_div_peephole(e, self.double_funs))
            e = _ternary_to_select_nested(e, self.double_funs)
            if zmul:
                e = _zmul_peephole(e)
        return _mangle_expr(e)

    def cse(self, e, cache, out, record=None, hinted=None, hin="odx_h"):
        """hoist sin(A) / cos(A) into one odx_sincos(A) per distinct argument (innermost calls first).
        record: {canonical argument: hint slot} -- after the sincos of such an argument, store (A, sin A, cos A) in *odx_h
        (the RHS side of the stage-to-stage increment); hinted: the same map on the %M side, where sincos(A') is then formed
        from the recorded triple and the small difference A' - A"""
        while True:
            found, skip = None, set()
            for fn in ("sin", "cos"):
                for b, en, args in _find_call(e, fn):
                    if len(args) == 1 and not re.search(r"\b(sin|cos)\s*\(", args[0]) and not (b, en) in skip:
                        if re.search(r"odx_c\.(rand01|randN|exprand|poisson|dwdt)\(", args[0]) or \
                                any(re.search(r"\b%s\s*\(" % re.escape(mangle(n)), args[0]) for n in self.rng_funs):
                            skip.add((b, en))        # a draw inside the argument: hoisting it out of a conditional arm would consume the draw
                            continue                 # even when the reference does not evaluate that arm; the call stays where it is
                        found = (b, en, fn, args[0])
                        break
                if found:
                    break
            if not found:
                return e
            b, en, fn, arg = found
            key = re.sub(r"\s+", "", arg)
            if key not in cache:
                cache[key] = len(cache)
                k = cache[key]
                canon = self.canon_arg(arg)
                if hinted is not None and canon in hinted:
                    j = hinted[canon]
                    out("    double odx_sn%d, odx_cs%d; odx_sincos_inc<ODX_EXACT>(odx_c, %s, %s.a[%d], %s.s[%d], %s.c[%d], &odx_sn%d, &odx_cs%d);"
                        % (k, k, arg, hin, j, hin, j, hin, j, k, k))
                else:
                    out("    double odx_sn%d, odx_cs%d; odx_sincos<ODX_EXACT>(odx_c, %s, &odx_sn%d, &odx_cs%d);" % (k, k, arg, k, k))
                if record is not None and canon in record:
                    j = record[canon]
                    out("    if (odx_h) { odx_h->a[%d] = %s; odx_h->s[%d] = odx_sn%d; odx_h->c[%d] = odx_cs%d; }" % (j, arg, j, k, j, k))
            k = cache[key]
            e = e[:b] + ("odx_sn%d" % k if fn == "sin" else "odx_cs%d" % k) + e[en:]

    def canon_arg(self, arg):
        """canonical text of a sincos argument in terms of odx_y[] / odx_e[] only, or None when it reads anything else
        (mean fields, auxiliaries, couplings, time, parameters, random numbers): only such arguments are functions of the
        particle state alone, so that the %M side of a stage sees the RHS side's argument shifted by a small step"""
        a = re.sub(r"\s+", "", arg)
        for e in self.m.eqs:
            a = re.sub(r"(?<![\w\]])%s(?![\w\[(])" % re.escape(mangle(e.name)), "odx_y[%d]" % self.var_idx[e.name], a)
        for e in self.m.exprs:
            if not e.length:
                a = re.sub(r"(?<![\w\]])%s(?![\w\[(])" % re.escape(mangle(e.name)), "odx_e[%d]" % self.expr_base[e.name], a)
        while True:
            b = re.sub(r"\((odx_[ye]\[\d+\])\)", r"\1", a)
            if b.startswith("(") and b.endswith(")"):
                depth, ok = 0, True
                for i, ch in enumerate(b):
                    depth += ch == "("
                    depth -= ch == ")"
                    if depth == 0 and i < len(b) - 1:
                        ok = False
                        break
                if ok:
                    b = b[1:-1]
            if b == a:
                break
            a = b
        rest = re.sub(r"odx_[ye]\[\d+\]|\d+\.?\d*(?:[eE][-+]?\d+)?|[-+*/()]", "", a)
        return a if rest == "" and "odx_y[" in a else None

    def sincos_args(self, exprs, where):
        """canonical sincos arguments the given expressions would hoist"""
        seen, lines = {}, []
        for x in exprs:
            if x:
                self.cse(self.tr(x, where), seen, lines.append)
        out = set()
        for ln in lines:
            mm = re.search(r"odx_sincos<ODX_EXACT>\(odx_c, (.*), &odx_sn\d+, &odx_cs\d+\);", ln)
            if mm and self.canon_arg(mm.group(1)):
                out.add(self.canon_arg(mm.group(1)))
        return out

    def uses(self, exprs, pat):
        return any(re.search(pat, x or "") for x in exprs)

    # ---- emit ------------------------------------------------------------------------------------
    def emit(self) -> str:
        m = self.m
        self._activate()
        pair_cpls = [c for c in m.cpls if c.lowrank is None]
        lr_cpls = [c for c in m.cpls if c.lowrank is not None]
        term_exprs = [e.rhs for e in m.mfds] + ([m.birth] if m.birth else [])
        cpl_exprs = [c.rhs for c in pair_cpls]
        rhs_exprs = [e.rhs for e in m.auxs] + [e.rhs for e in m.eqs] + [m.death, m.repli]
        if self.uses(term_exprs, r"\b(OA|SA)\(") or self.uses(cpl_exprs, r"\bSA\("):
            raise PopError("OA()/SA() inside %M and SA() inside %C are not supported")
        # OA(a) inside %C: in the reference other_->aux holds pass-2 values for the particles before myself_ in the list and pass-1
        # values (zeroed mean fields, previous psi) for those after it (src/odexp:905-954).  The two coincide -- and the sum can be
        # formed in parallel -- exactly when `a`, through the auxiliaries it reads, depends on no mean field and no coupling.
        cpl_uses_oaux = self.uses(cpl_exprs, r"\bOA\(")
        if cpl_uses_oaux:
            tainted = {x.name for x in m.mfds} | {x.name for x in m.cpls}
            dirty = set()
            for a in m.auxs:                                   # file order: an auxiliary can only read the ones before it
                ids = set(re.findall(IDENT, a.rhs))
                if ids & (tainted | dirty) or re.search(r"\b(MF|MC|KY)\b", a.rhs) or re.search(self.draw_pat, a.rhs):
                    dirty.add(a.name)
            for ex in cpl_exprs:
                for nm in re.findall(r"\bOA\(\s*(%s)\s*\)" % IDENT, ex or ""):
                    if nm in dirty:
                        raise PopError("OA(%s) inside %%C: '%s' depends on a mean field, a coupling or a random number, so its value for "
                                       "the other particle depends on the list order in the reference; not supported" % (nm, nm))
        cpl_uses_aux = self.uses(cpl_exprs, r"\bMA\(")       # the particle's own pass-1 auxiliary values (src/odexp:820-833, 905-954)
        if self.uses(term_exprs, r"\b(OY|OE)\("):
            raise PopError("OY()/OE() are only meaningful inside %C")
        if self.uses(term_exprs + cpl_exprs, self.draw_pat):
            # the reference draws from the serial libc stream once per particle (%M, %BIRTH) or per ordered pair (%C) in list order; the
            # population sums here run in parallel in another order, and their contexts carry no random stream
            raise PopError("random numbers inside %M, %C or %BIRTH expressions are not supported (draw them in an EXPR, AUX or D/DT "
                           "expression and read that through ME() / MA())")
        terms_use_aux = self.uses(term_exprs, r"\bMA\(")
        uses_ky = self.uses(rhs_exprs, r"\bKY\s*\(")
        if self.uses(term_exprs + cpl_exprs + [x.rhs for x in m.exprs] + [x.rhs for x in m.inits], r"\bKY\s*\("):
            raise PopError("KY() ring neighbours are only supported inside AUX / D/DT / %DEATH / %REPLI")
        if uses_ky and terms_use_aux:
            raise PopError("KY() together with MA() inside %M is not supported (the reference's first pass sees a half-updated list)")
        cpl_uses_expr = self.uses(cpl_exprs, r"\bOE\(") or (cpl_uses_oaux and m.nexpr > 0)      # the other particle's aux may read its expressions
        rhs_rng = self.uses(rhs_exprs, self.draw_pat)
        if cpl_uses_aux and rhs_rng:
            raise PopError("MA() inside %C in a model with random numbers in its RHS is not supported (the pass-1 evaluation would consume draws)")
        has_stopif = self.uses(rhs_exprs, r"\bSTOPIF\s*\(")
        # the particle id column is only read by the stage kernels when something can observe it
        needs_id = rhs_rng or self.uses(rhs_exprs + term_exprs + cpl_exprs, r"\b(ID|OID|MID|rand01|randN|exprand|poisson|DWDT)\b")
        L = []
        w = L.append
        w("/* generated by odexp_b200.cudagen from model '%s' — do not edit */" % m.name)
        w('#include <math.h>\n#include <stdint.h>\n#include "odx.h"\n#include "odx_rng.h"\n#include "odx_model_prelude.cuh"\n')
        for name, val in m.macros:
            w("#define %s %s" % (mangle(name), _mangle_expr(val)))
        # constant scalars / arrays (`decl const double a[] = {..}`, file scope in the reference's model.c): constexpr locals of every
        # generated function, so that host and device code index the same values without a __constant__ / host copy pair
        w("#define ODX_CONSTS %s" % " ".join("constexpr %s %s%s = %s; (void)%s;" % (t, mangle(n), d, _mangle_expr(init), mangle(n))
                                             for t, n, d, init in m.consts))
        w("")
        for f in m.funs:
            # arguments are doubles unless they carry a type; `*x` passes an array (src/odexp:694-707)
            args = ", ".join(_mangle_expr(a) if re.search(r"\b(void|int|long|char|float|const)\b", a) else
                             "double " + _mangle_expr(a) for a in f.args)
            if f.name in self.ctx_funs:
                # MU(par) / POP_SIZE / random draws inside the function (the reference reaches them through globals): the evaluation
                # context comes in as the first argument, every call site passes its own (tr / _ctx_calls); no arithmetic rewrites
                rt = re.search(r"return\s*\((int|long|char|float)\)", f.body) if f.block else None
                inner = _mangle_expr(self._fun_body(f, f.body))
                w("template <class OdxC> ODX_HD %s %s(OdxC &odx_c%s) { (void)odx_c; ODX_CONSTS %s }" %
                  (rt.group(1) if rt else "double", mangle(f.name), ", " + args if args else "", inner if f.block else "return %s;" % inner))
                continue
            if f.block:
                rt = re.search(r"return\s*\((int|long|char|float)\)", f.body)
                w("ODX_HD %s %s(%s) { ODX_CONSTS %s }" % (rt.group(1) if rt else "double", mangle(f.name), args, _mangle_expr(f.body)))
                continue
            body = _pow_peephole(f.body) if self.fast_pow else f.body
            body = _div_peephole(body, self.double_funs) if self.fast_div else body
            body = _mangle_expr(_ternary_to_select(body, self.double_funs))
            w("ODX_HD double %s(%s) { ODX_CONSTS return %s; }" % (mangle(f.name), args, body))
        w("")
        # sincos arguments shared by the RHS and the %M terms: candidates for the stage-to-stage increment
        hint_slots = {}
        if not terms_use_aux and not rhs_rng:
            shared = self.sincos_args([e.rhs for e in m.auxs if not e.length] + [e.rhs for e in m.eqs if not e.length], "rhs") & \
                     self.sincos_args([e.rhs for e in m.mfds if not is_vect(e)], "%M")
            hint_slots = {a: i for i, a in enumerate(sorted(shared))}
        self.hint_slots = hint_slots
        w("/* (argument, sin, cos) of the sincos arguments the RHS shares with the %M terms */")
        w("#define ODX_NHINT %d" % len(hint_slots))
        w("struct OdxHint { double a[%d], s[%d], c[%d]; };" % ((max(1, len(hint_slots)),) * 3))
        w("struct OdxModel {")
        w("  static constexpr int NVAR = %d, NEXPR = %d, NAUX = %d, NPSI = %d, NMFD = %d, NPAR = %d;" %
          (m.nvar, m.nexpr, m.naux, m.npsi, m.nmfd, m.npar))
        w("  static constexpr bool HAS_BIRTH = %s, HAS_DEATH = %s, HAS_REPLI = %s, RHS_RNG = %s;" %
          tuple("true" if x else "false" for x in (m.birth, m.death, m.repli, rhs_rng)))
        w("  static constexpr bool HAS_STOPIF = %s;" % ("true" if has_stopif else "false"))
        w("  static constexpr bool CPL_USES_AUX = %s;      /* MA() inside %%C: ode_rhs with zeroed mean fields and the previous psi first */" % ("true" if cpl_uses_aux else "false"))
        w("  static constexpr bool CPL_USES_OAUX = %s;     /* OA() inside %%C: the other particle's auxiliaries, evaluated per staged j-particle */" % ("true" if cpl_uses_oaux else "false"))
        w("  static constexpr bool USES_KY = %s;      /* KY(name, degree): the RHS reads the stage state of list neighbours */" % ("true" if uses_ky else "false"))
        w("  static constexpr int NHINT = ODX_NHINT;")
        w("  static constexpr bool HAS_PAIR = %s, HAS_LOWRANK = %s, TERMS_USE_AUX = %s, CPL_USES_EXPR = %s, NEEDS_ID = %s;" %
          tuple("true" if x else "false" for x in (pair_cpls, lr_cpls, terms_use_aux, cpl_uses_expr, needs_id)))
        w("")
        # ---- ode_rhs (and the variant whose shared sincos arguments come from the previous stage's triple)
        for inc in (False, True):
            if inc:
                w("  /* the RHS at a state one stage after the one described by odx_hi (the lean RK4 scheme recomputes stages) */")
            w("  /* per-particle right-hand side: AUX, D/DT, %DEATH, %REPLI (reference: generated ode_rhs) */")
            w("  /* ODX_EXACT = false: range-limited fast sincos, the context records when an argument was out of range and the engine")
            w("   * redoes the particle with ODX_EXACT = true (library sincos); keeps the hot code free of branches */")
            w("  template <bool ODX_EXACT, class C> ODX_HD static void %s(C &odx_c, double t, const double *odx_y, const double *odx_e," % ("rhs_inc" if inc else "rhs"))
            w("      const double *odx_psi, const double *odx_mfd, double *odx_f, double *odx_aux, double &odx_death, double &odx_repli,")
            w("      %sOdxHint *odx_h = nullptr) {" % ("const OdxHint &odx_hi, " if inc else ""))
            w("    (void)t; (void)odx_psi; (void)odx_mfd; (void)odx_e; (void)odx_h;%s" % (" (void)odx_hi;" if inc else ""))
            w("    ODX_CONSTS")
            for i, p in enumerate(m.pars):
                if not re.search(r"\b(init|impl)", p.attr, flags=re.I):
                    ctype = "double"
                    mt = re.search(r"\b(int|long|double)\b", p.attr)
                    if mt:
                        ctype = mt.group(1)
                    w("    const %s %s = (%s)odx_c.mu[%d];" % (ctype, mangle(p.name), ctype, i))
            for e in m.exprs:
                b = self.expr_base[e.name]
                if _nd(e):
                    w("    double %s; %s %s = odx_e[%d + %s];" % (_arr_decl(e), _arr_loops(e), _arr_ref(e), b, _arr_flat(e)))
                elif e.length:
                    w("    double %s[%d]; for (int odx_i = 0; odx_i < %d; odx_i++) %s[%d + odx_i] = odx_e[%d + odx_i];" %
                      (mangle(e.name), e.start + e.length, e.length, mangle(e.name), e.start, b))
                else:
                    w("    const double %s = odx_e[%d];" % (mangle(e.name), b))
            for e in m.eqs:
                if _nd(e):
                    w("    double %s; %s %s = odx_y[%d + %s];" % (_arr_decl(e), _arr_loops(e), _arr_ref(e), self.var_idx[e.name], _arr_flat(e)))
                elif e.length:
                    w("    double %s[%d]; for (int odx_i = 0; odx_i < %d; odx_i++) %s[%d + odx_i] = odx_y[%d + odx_i];" %
                      (mangle(e.name), e.start + e.length, e.length, mangle(e.name), e.start, self.var_idx[e.name]))
                else:
                    w("    const double %s = odx_y[%d];" % (mangle(e.name), self.var_idx[e.name]))
            for c in m.cpls:
                w("    const double %s = odx_psi[%d];" % (mangle(c.name), self.psi_idx[c.name]))
            for e in m.mfds:
                w("    const double %s = odx_mfd[%d];" % (mangle(e.name), self.mfd_idx[e.name]))
            tc = {}
            # arrays: one loop per entry with the reference's `size_t` iterator (src/odexp:1381-1436, 1439-1506); no common-subexpression
            # hoisting across the loop
            for e in m.auxs:
                if _nd(e) and is_vect(e):
                    w("    double %s; %s;" % (_arr_decl(e), self.tr(e.rhs, "aux " + e.name)))
                elif _nd(e):
                    w("    double %s; %s %s = %s;" % (_arr_decl(e), _arr_loops(e), _arr_ref(e), self.tr(e.rhs, "aux " + e.name)))
                elif e.length and is_vect(e):        # filled by a function: the rhs is evaluated once, as a statement (src/odexp:1394-1397)
                    w("    double %s[%d]; %s;" % (mangle(e.name), e.start + e.length, self.tr(e.rhs, "aux " + e.name)))
                elif e.length:
                    it = mangle(e.iterator)
                    w("    double %s[%d]; for (size_t %s = %d; %s < %d; ++%s) %s[%s] = %s;" % (mangle(e.name), e.start + e.length, it, e.start, it,
                      e.start + e.length, it, mangle(e.name), it, self.tr(e.rhs, "aux " + e.name)))
                else:
                    w("    const double %s = %s;" % (mangle(e.name), self.cse(self.tr(e.rhs, "aux " + e.name), tc, w, record=hint_slots, hinted=(hint_slots if inc else None), hin="odx_hi")))
            for e in m.eqs:
                if _nd(e):
                    w("    %s odx_f[%d + %s] = %s;" % (_arr_loops(e), self.var_idx[e.name], _arr_flat(e), self.tr(e.rhs, "d%s/dt" % e.name)))
                elif e.length:
                    it = mangle(e.iterator)
                    w("    for (size_t %s = %d; %s < %d; ++%s) odx_f[%d + %s - %d] = %s;" % (it, e.start, it, e.start + e.length, it,
                      self.var_idx[e.name], it, e.start, self.tr(e.rhs, "d%s/dt" % e.name)))
                else:
                    w("    odx_f[%d] = %s;" % (self.var_idx[e.name], self.cse(self.tr(e.rhs, "d%s/dt" % e.name), tc, w, record=hint_slots, hinted=(hint_slots if inc else None), hin="odx_hi")))
            for e in m.auxs:
                if _nd(e):
                    w("    %s odx_aux[%d + %s] = %s;" % (_arr_loops(e), self.aux_idx[e.name], _arr_flat(e), _arr_ref(e)))
                elif e.length:
                    w("    for (int odx_i = 0; odx_i < %d; odx_i++) odx_aux[%d + odx_i] = %s[%d + odx_i];" % (e.length, self.aux_idx[e.name], mangle(e.name), e.start))
                else:
                    w("    odx_aux[%d] = %s;" % (self.aux_idx[e.name], mangle(e.name)))
            if m.death:
                w("    odx_death = %s;" % self.cse(self.tr(m.death, "%DEATH"), tc, w))
            if m.repli:
                w("    odx_repli = %s;" % self.cse(self.tr(m.repli, "%REPLI"), tc, w))
            w("  }\n")
        # ---- %M terms
        w("  /* per-particle %M expressions (summed over the population, then divided by POP_SIZE) */")
        w("  template <bool ODX_EXACT, class C> ODX_HD static void mfd_terms(C &odx_c, double t, const double *odx_y, const double *odx_e,")
        w("      const double *odx_aux, double *odx_m) {")
        w("    (void)t; (void)odx_y; (void)odx_e; (void)odx_aux; (void)odx_c; (void)odx_m;")
        w("    ODX_CONSTS")
        tc = {}
        for e in m.mfds:
            if is_vect(e):
                w("    odx_m[%d] = 0.0;      /* {vect}: no per-particle term, see mfd_vect */" % self.mfd_idx[e.name])
                continue
            w("    odx_m[%d] = %s;" % (self.mfd_idx[e.name], self.cse(self.tr(e.rhs, "%M " + e.name), tc, w)))
        w("  }\n")
        w("  /* the same terms for the state one stage after the one the RHS was evaluated at: sin / cos of an argument the RHS has")
        w("   * already reduced come from its (argument, sin, cos) triple and the small difference (angle addition) */")
        w("  template <bool ODX_EXACT, class C> ODX_HD static void mfd_terms_inc(C &odx_c, double t, const double *odx_y, const double *odx_e,")
        w("      const double *odx_aux, double *odx_m, const OdxHint &odx_h) {")
        w("    (void)t; (void)odx_y; (void)odx_e; (void)odx_aux; (void)odx_c; (void)odx_m; (void)odx_h;")
        w("    ODX_CONSTS")
        tc = {}
        for e in m.mfds:
            if is_vect(e):
                w("    odx_m[%d] = 0.0;" % self.mfd_idx[e.name])
                continue
            w("    odx_m[%d] = %s;" % (self.mfd_idx[e.name], self.cse(self.tr(e.rhs, "%M " + e.name), tc, w, hinted=hint_slots)))
        w("  }\n")
        # ---- %M ... {vect}
        vects = [e for e in m.mfds if is_vect(e)]
        w("  /* %M ... {vect}: evaluated once per RHS evaluation from POP_SIZE, the parameters and the mean fields before it, in file order")
        w("   * (reference: `SIM->meanfield[k] = rhs;` instead of the list walk, src/odexp:886-889) */")
        w("  static constexpr bool HAS_VECT_MFD = %s;" % ("true" if vects else "false"))
        w("  ODX_HD static void mfd_vect(const double *odx_mu, double odx_pop_size, double *odx_mfd) {")
        w("    (void)odx_mu; (void)odx_pop_size; (void)odx_mfd;")
        w("    ODX_CONSTS")
        if vects:
            w("    struct { const double *mu; double pop_size; } odx_c = {odx_mu, odx_pop_size}; (void)odx_c;")
        for e in vects:
            if re.search(r"\b(MY|ME|MA|OY|OE|OA|SY|SE|SA|MC|KY|DWDT|rand01|randN)\b", e.rhs):
                raise PopError("%%M %s {vect}: the expression is evaluated once, it cannot read a particle (%s)" % (e.name, e.rhs))
            w("    odx_mfd[%d] = %s;" % (self.mfd_idx[e.name], self.tr(e.rhs, "%M " + e.name + " {vect}", exact=True)))
        w("  }\n")
        w("  template <bool ODX_EXACT, class C> ODX_HD static double birth_term(C &odx_c, double t, const double *odx_y, const double *odx_e,")
        w("      const double *odx_aux) {")
        w("    (void)t; (void)odx_y; (void)odx_e; (void)odx_aux; (void)odx_c;")
        w("    ODX_CONSTS")
        w("    return %s;" % (self.tr(m.birth, "%BIRTH") if m.birth else "0.0"))
        w("  }\n")
        # ---- %C terms
        w("  /* pairwise %C expressions: odx_p[k] += expr_k(me, other) */")
        w("  template <bool ODX_EXACT, class C> ODX_HD static void cpl_terms(C &odx_c, double t, const double *odx_y, const double *odx_e,")
        w("      const double *odx_aux, const double *odx_oy, const double *odx_oe, const double *odx_oaux, double *odx_p) {")
        w("    (void)t; (void)odx_y; (void)odx_e; (void)odx_aux; (void)odx_oy; (void)odx_oe; (void)odx_oaux; (void)odx_c; (void)odx_p;")
        w("    ODX_CONSTS")
        tc = {}
        scales = {}
        sym_terms = []
        for c in pair_cpls:
            term = self.tr(c.rhs, "%C " + c.name, zmul=self.fast_div)
            if self.fast_div:
                scales[c.name], term = _hoist_invariant_factors(term)
                term = _strip_outer_parens(term)
            sym_terms.append((self.psi_idx[c.name], _antisymmetric_pair_term(term)))
            w("    odx_p[%d] += %s;" % (self.psi_idx[c.name], self.cse(term, tc, w)))
        w("  }\n")
        # ---- antisymmetric couplings (Newton's third law): one evaluation per unordered pair
        self.cpl_sym = bool(pair_cpls) and not lr_cpls and self.fast_div and all(st is not None for _, st in sym_terms) \
            and not (cpl_uses_aux or cpl_uses_oaux)
        w("  /* every %C term changes sign (bit for bit) when the two particles are exchanged: k_pair_sym evaluates each unordered pair once */")
        w("  static constexpr bool CPL_SYM = %s;" % ("true" if self.cpl_sym else "false"))
        if self.cpl_sym:
            ymask = 0
            for _, (k, R, used) in sym_terms:
                for u in used:
                    ymask |= 1 << u
            w("  static constexpr unsigned CPL_SYM_YMASK = 0x%xu;      /* dynamical variables the terms read */" % ymask)
            w("  template <bool ODX_EXACT, class C> ODX_HD static void cpl_terms_sym(C &odx_c, double t, const double *odx_y, const double *odx_oy,")
            w("      double *odx_p, double *odx_q) {")
            w("    (void)t; (void)odx_c;")
            w("    ODX_CONSTS")
            for pi, (k, R, used) in sym_terms:
                w("    { const double odx_d = odx_y[%d] - odx_oy[%d], odx_r = %s;" % (k, k, R.strip()))
                w("      odx_p[%d] = fma(odx_d, odx_r, odx_p[%d]); odx_q[%d] = fma(-odx_d, odx_r, odx_q[%d]); }" % (pi, pi, pi, pi))
            w("  }\n")
        w("  /* factors of the %C terms that do not depend on the other particle: psi_k(i) = scale_k(i) * sum_j term_k(i,j) */")
        w("  static constexpr bool CPL_SCALED = %s;" % ("true" if any(scales.values()) else "false"))
        w("  /* the fast %C terms signal an unrepresentable term by a non-finite sum (odx_zmul) */")
        w("  static constexpr bool CPL_POISON = %s;" % ("true" if any("odx_zmul<" in ln for ln in L) else "false"))
        w("  template <bool ODX_EXACT, class C> ODX_HD static void cpl_scale(C &odx_c, double t, const double *odx_y, const double *odx_e,")
        w("      const double *odx_aux, double *odx_s) {")
        w("    (void)t; (void)odx_y; (void)odx_e; (void)odx_aux; (void)odx_c; (void)odx_s;")
        w("    ODX_CONSTS")
        for c in pair_cpls:
            w("    odx_s[%d] = %s;" % (self.psi_idx[c.name], scales.get(c.name) or "1.0"))
        w("  }\n")
        # ---- initial conditions
        # ---- delay() call sites: what the engine needs to carry out their side effects between evaluations
        nd = len(self.delay_sites)
        if nd and terms_use_aux:
            raise PopError("delay() in a model whose %M terms read an auxiliary variable through MA() is not supported (the reference's first pass "
                           "would look the history up before / after the push of the step)")
        if nd and (m.birth or m.death or m.repli):
            raise PopError("delay() together with %BIRTH / %DEATH / %REPLI is not supported (the history stacks are indexed by particle)")
        w("  /* delay(t, x, tau, ic, &stack[idx]) call sites (src/utils_odexp.c:205-240): (time, value) pushed at the first evaluation of a dde")
        w("   * step, the stack the particle uses, and the history function that fills the stack at t0 */")
        w("  static constexpr int NDELAY = %d;" % nd)
        w("  ODX_HD static constexpr long delay_count(int k) { return %s; }" %
          (" : ".join(["k == %d ? %dL" % (k, st["count"]) for k, st in enumerate(self.delay_sites)] + ["1L"]) if nd else "1L"))
        w("  template <class C> ODX_HD static void delay_sample(C &odx_c, int odx_k, double t, const double *odx_y, const double *odx_e,")
        w("      double &odx_tt, double &odx_xx, long &odx_idx) {")
        w("    (void)odx_c; (void)odx_k; (void)t; (void)odx_y; (void)odx_e; odx_tt = t; odx_xx = 0.0; odx_idx = 0;")
        w("    ODX_CONSTS")
        if nd:
            for i, p in enumerate(m.pars):
                if not re.search(r"\b(init|impl)", p.attr, flags=re.I):
                    ctype = "double"
                    mt = re.search(r"\b(int|long|double)\b", p.attr)
                    if mt:
                        ctype = mt.group(1)
                    w("    const %s %s = (%s)odx_c.mu[%d]; (void)%s;" % (ctype, mangle(p.name), ctype, i, mangle(p.name)))
            for e in m.exprs:
                b = self.expr_base[e.name]
                if _nd(e):
                    w("    double %s; %s %s = odx_e[%d + %s];" % (_arr_decl(e), _arr_loops(e), _arr_ref(e), b, _arr_flat(e)))
                elif e.length:
                    w("    double %s[%d]; for (int odx_i = 0; odx_i < %d; odx_i++) %s[%d + odx_i] = odx_e[%d + odx_i];" %
                      (mangle(e.name), e.start + e.length, e.length, mangle(e.name), e.start, b))
                else:
                    w("    const double %s = odx_e[%d]; (void)%s;" % (mangle(e.name), b, mangle(e.name)))
            for e in m.eqs:
                if _nd(e):
                    w("    double %s; %s %s = odx_y[%d + %s];" % (_arr_decl(e), _arr_loops(e), _arr_ref(e), self.var_idx[e.name], _arr_flat(e)))
                elif e.length:
                    w("    double %s[%d]; for (int odx_i = 0; odx_i < %d; odx_i++) %s[%d + odx_i] = odx_y[%d + odx_i];" %
                      (mangle(e.name), e.start + e.length, e.length, mangle(e.name), e.start, self.var_idx[e.name]))
                else:
                    w("    const double %s = odx_y[%d]; (void)%s;" % (mangle(e.name), self.var_idx[e.name], mangle(e.name)))
            for k, st in enumerate(self.delay_sites):
                w("    if (odx_k == %d) { odx_tt = %s; odx_xx = %s; odx_idx = (long)(%s); }" %
                  (k, self.tr(st["t"], "delay", exact=True), self.tr(st["x"], "delay", exact=True), self.tr(st["idx"], "delay", exact=True)))
        w("  }")
        w("  ODX_HD static double delay_ic(int odx_k, double odx_s) {")
        w("    (void)odx_k; (void)odx_s;")
        for k, st in enumerate(self.delay_sites):
            w("    if (odx_k == %d) return %s(odx_s);" % (k, mangle(st["ic"])))
        w("    return 0.0;")
        w("  }\n")
        w("  /* ode_init_conditions: EXPR in file order, then INIT in file order (draw order matters for replay) */")
        w("  template <class C> ODX_HD static void ic(C &odx_c, double t, double *odx_y, double *odx_e) {")
        w("    (void)t; (void)odx_e;")
        w("    ODX_CONSTS")
        for i, p in enumerate(m.pars):
            if re.search(r"\b(init|ever)", p.attr, flags=re.I):
                w("    const double %s = odx_c.mu[%d];" % (mangle(p.name), i))
        for e in m.exprs:
            b = self.expr_base[e.name]
            if _nd(e) and is_vect(e):
                w("    double %s; %s;" % (_arr_decl(e), self.tr(e.rhs, "expr " + e.name, exact=True)))
                w("    %s odx_e[%d + %s] = %s;" % (_arr_loops(e), b, _arr_flat(e), _arr_ref(e)))
            elif _nd(e):
                w("    double %s;" % _arr_decl(e))
                w("    %s { %s = %s; odx_e[%d + %s] = %s; }" % (_arr_loops(e), _arr_ref(e), self.tr(e.rhs, "expr " + e.name, exact=True), b, _arr_flat(e), _arr_ref(e)))
            elif e.length and is_vect(e):
                w("    double %s[%d]; %s;" % (mangle(e.name), e.start + e.length, self.tr(e.rhs, "expr " + e.name, exact=True)))
                w("    for (int odx_i = 0; odx_i < %d; odx_i++) odx_e[%d + odx_i] = %s[%d + odx_i];" % (e.length, b, mangle(e.name), e.start))
            elif e.length:
                it = mangle(e.iterator)
                w("    double %s[%d];" % (mangle(e.name), e.start + e.length))
                w("    for (size_t %s = %d; %s < %d; %s++) { %s[%s] = %s; odx_e[%d + %s - %d] = %s[%s]; }" %
                  (it, e.start, it, e.start + e.length, it, mangle(e.name), it, self.tr(e.rhs, "expr " + e.name, exact=True), b, it, e.start,
                   mangle(e.name), it))
            else:
                w("    const double %s = %s; odx_e[%d] = %s;" % (mangle(e.name), self.tr(e.rhs, "expr " + e.name, exact=True), b, mangle(e.name)))
        for e in m.inits:
            if _nd(e):
                w("    %s odx_y[%d + %s] = %s;" % (_arr_loops(e), self.var_idx[e.name], _arr_flat(e), self.tr(e.rhs, "init " + e.name, exact=True)))
            elif e.length:
                it = mangle(e.iterator)
                w("    for (size_t %s = %d; %s < %d; ++%s) odx_y[%d + %s - %d] = %s;" % (it, e.start, it, e.start + e.length, it, self.var_idx[e.name],
                  it, e.start, self.tr(e.rhs, "init " + e.name, exact=True)))
            else:
                w("    odx_y[%d] = %s;" % (self.var_idx[e.name], self.tr(e.rhs, "init " + e.name, exact=True)))
        w("  }\n")
        w("  ODX_HD static constexpr bool psi_is_pairwise(int k) { return %s; }\n" %
          (" || ".join("k == %d" % self.psi_idx[c.name] for c in pair_cpls) or "false"))
        # ---- low-rank coupling descriptors
        w("  static constexpr int NLOWRANK = %d;" % len(lr_cpls))
        for k, c in enumerate(lr_cpls):
            d = c.lowrank
            var = d.get("var")
            # the generated gather uses mvar(): dynamical variables, then auxiliary variables, then expressions (src/datastruct.c:824-855)
            if var in self.var_idx:
                var_kind, var_i = 0, self.var_idx[var]
            elif var in self.aux_idx:
                var_kind, var_i = 1, self.aux_idx[var]
                if rhs_rng:
                    raise PopError("low-rank coupling %s: var=%s is an auxiliary variable of a model with random numbers in its RHS "
                                   "(its pass-1 value would consume draws)" % (c.name, var))
            elif var in self.expr_base and not next(e for e in m.exprs if e.name == var).length:
                var_kind, var_i = 2, self.expr_base[var]
            else:
                raise PopError("low-rank coupling %s: var=%s is not a dynamical variable, an auxiliary variable or a scalar expression" % (c.name, var))
            fun = d.get("fun", "sin")
            rank = int(d.get("rank", "1")) if d["kind"] == "lrwexp" else 1          # lrexp = rank 1 with unit weights (lrkern, src/lrexp.c:484-496)
            # `U = U[0]` (the shipped examples/kuramoto_lowrank.pop) names the first element of the array: same column base
            base = lambda key: self.expr_base.get(re.sub(r"\[\s*0\s*\]$", "", d.get(key, "").strip()), -1)
            ub = base("U") if d["kind"] == "lrwexp" else -1
            vb = base("V") if d["kind"] == "lrwexp" else -1
            if d["kind"] == "lrwexp" and (ub < 0 or vb < 0):
                raise PopError("low-rank coupling %s: U/V must name EXPR arrays" % c.name)
            w("  static constexpr int LR%d_PSI = %d, LR%d_VAR = %d, LR%d_RANK = %d, LR%d_U = %d, LR%d_V = %d;" %
              (k, self.psi_idx[c.name], k, var_i, k, rank, k, ub, k, vb))
            w("  static constexpr bool LR%d_UNWEIGHTED = %s;      /* %%C ... lrexp: U = V = 1, no EXPR columns are read */" % (k, "true" if d["kind"] != "lrwexp" else "false"))
            w("  static constexpr int LR%d_VAR_KIND = %d;      /* LR%d_VAR indexes 0: Y, 1: AUX (pass-1 value), 2: EXPR */" % (k, var_kind, k))
            w("  ODX_HD static double lr%d_fun(double x) { return %s(x); }" % (k, mangle(fun)))
        w("")
        # ---- static info
        def cstr_list(names):
            return "{" + ", ".join('"%s"' % n.replace('"', "") for n in names) + (", " if names else "") + "nullptr}"
        w("  static const odx_model_info &info() {")
        w("    static const char *const vn[] = %s;" % cstr_list(m.flat_names(m.eqs)))
        w("    static const char *const en[] = %s;" % cstr_list(m.flat_names(m.exprs)))
        w("    static const char *const an[] = %s;" % cstr_list(m.flat_names(m.auxs)))
        w("    static const char *const pn[] = %s;" % cstr_list([c.name for c in m.cpls]))
        w("    static const char *const mn[] = %s;" % cstr_list([e.name for e in m.mfds]))
        w("    static const char *const qn[] = %s;" % cstr_list([e.name for e in m.pars]))
        vals = ", ".join(repr(float(p.rhs)) for p in m.pars) or "0.0"
        w("    static const double pd[] = {%s};" % vals)
        w('    static const odx_model_info mi = {"%s", NVAR, NEXPR, NAUX, NPSI, NMFD, NPAR, HAS_BIRTH, HAS_DEATH, HAS_REPLI, RHS_RNG,' % m.name)
        w("                                      HAS_PAIR, HAS_LOWRANK, vn, en, an, pn, mn, qn, pd, HAS_STOPIF, NDELAY};")
        w("    return mi;")
        w("  }")
        w("};\n")
        if lr_cpls:
            if len(lr_cpls) > 1:
                raise PopError("only one low-rank coupling per model is supported")
            w("#define ODX_MODEL_HAS_LOWRANK 1")
        if getattr(self, "cpl_sym", False):
            w("#define ODX_MODEL_CPL_SYM 1")
        if m.macros:
            # the model's MACRO lines are #defines, as in the reference's model.c -- but there they live in a translation unit of their own.
            # Here the engine follows in the same one: a model's `macro T 1.0`, `macro h ...` or `macro n ...` would rewrite the engine's own
            # identifiers.  The generated functions above are already preprocessed text, so the macros can end here.
            w("/* end of the model's macros (the engine below must not see them) */")
            for name, _ in m.macros:
                w("#undef %s" % mangle(name))
        w('#include "odx_engine.cuh"')
        text = "\n".join(L) + "\n"
        if re.search(r"\bodx_(max|min)\s*\(", text):
            text = text.replace('#include "odx_model_prelude.cuh"\n', '#include "odx_model_prelude.cuh"\n' + _MINMAX, 1)
        if re.search(r"\bodx_iabs\s*\(", text):
            text = text.replace('#include "odx_model_prelude.cuh"\n', '#include "odx_model_prelude.cuh"\n'
                                'ODX_HD int odx_iabs(int v) { return v < 0 ? -v : v; }      /* C\'s abs(): the argument converts to int */\n', 1)
        return text


def emit_cuda(model: PopModel, fast_pow: bool = True, fast_div: bool = True) -> str:
    return Emitter(model, fast_pow, fast_div).emit()
