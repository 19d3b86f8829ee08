"""Evaluator for the subset of OpenMM's Lepton expression language used by ``alchemical_functions``.

The reference hands strings such as ``'min(1, (1/0.3)*abs(lambda-0.5))'`` to the integrator
(``blues/simulation.py:654-659``), which OpenMM compiles with Lepton.  Here each expression is parsed once
into a Python AST restricted to arithmetic, the Lepton function set and the variable ``lambda``; the host
then tabulates it at every ``lambda_step`` so the device only ever indexes a table.
"""
import ast
import math
import re

_FUNCS = {
    'min': min, 'max': max, 'abs': abs, 'sqrt': math.sqrt, 'exp': math.exp, 'log': math.log, 'sin': math.sin,
    'cos': math.cos, 'tan': math.tan, 'asin': math.asin, 'acos': math.acos, 'atan': math.atan, 'sinh': math.sinh,
    'cosh': math.cosh, 'tanh': math.tanh, 'erf': math.erf, 'erfc': math.erfc, 'floor': math.floor,
    'ceil': math.ceil, 'step': lambda x: 1.0 if x >= 0 else 0.0, 'delta': lambda x: 1.0 if x == 0 else 0.0,
    'select': lambda c, a, b: a if c != 0 else b, 'square': lambda x: x * x, 'cube': lambda x: x * x * x,
    'recip': lambda x: 1.0 / x,
}
_BINOPS = {ast.Add: lambda a, b: a + b, ast.Sub: lambda a, b: a - b, ast.Mult: lambda a, b: a * b,
           ast.Div: lambda a, b: a / b, ast.Pow: lambda a, b: a ** b}
_VAR = '__lambda__'


class Discrete1DFunction(object):
    """OpenMM's tabulated function of an integer argument: f(x) = values[round(x)] (``CustomIntegrator.addTabulatedFunction``;
    what ``blues/utils.py:276-369`` ``spreadLambdaProtocol`` returns)."""

    def __init__(self, values):
        self.values = [float(v) for v in values]
        if not self.values:
            raise ValueError('a tabulated function needs at least one value')

    def getFunctionParameters(self):
        return list(self.values)

    def setFunctionParameters(self, values):
        self.values = [float(v) for v in values]

    def __call__(self, x):
        i = int(round(x))
        if i < 0 or i >= len(self.values):
            return 0.0                       # OpenMM: zero outside the tabulated range
        return self.values[i]


class Continuous1DFunction(object):
    """OpenMM's natural cubic spline through ``values`` at evenly spaced points of [min, max]; zero outside."""

    def __init__(self, values, min, max):
        import numpy as np
        from scipy.interpolate import CubicSpline
        self.values, self.min, self.max = [float(v) for v in values], float(min), float(max)
        if len(self.values) < 2 or not self.max > self.min:
            raise ValueError('a continuous tabulated function needs >= 2 values and max > min')
        self._spline = CubicSpline(np.linspace(self.min, self.max, len(self.values)), self.values, bc_type='natural')

    def getFunctionParameters(self):
        return list(self.values), self.min, self.max

    def __call__(self, x):
        if x < self.min or x > self.max:
            return 0.0
        return float(self._spline(x))


class Expression(object):
    """``functions``: name -> callable of one argument (tabulated functions added to the integrator)."""

    def __init__(self, text, functions=None):
        self.text = str(text)
        self._funcs = dict(_FUNCS)
        self._funcs.update(functions or {})
        src = self.text.split(';')[0].replace('^', '**')
        src = re.sub(r'\blambda\b', _VAR, src)
        self._tree = ast.parse(src.strip(), mode='eval').body
        self._check(self._tree)

    def _check(self, node):
        if isinstance(node, ast.BinOp) and type(node.op) in _BINOPS:
            self._check(node.left)
            self._check(node.right)
        elif isinstance(node, ast.UnaryOp) and isinstance(node.op, (ast.USub, ast.UAdd)):
            self._check(node.operand)
        elif isinstance(node, ast.Call) and isinstance(node.func, ast.Name) and node.func.id in self._funcs and not node.keywords:
            for a in node.args:
                self._check(a)
        elif isinstance(node, ast.Constant) and isinstance(node.value, (int, float)):
            pass
        elif isinstance(node, ast.Name) and node.id == _VAR:
            pass
        else:
            raise ValueError('unsupported element in expression %r' % self.text)

    def _eval(self, node, lam):
        if isinstance(node, ast.BinOp):
            return _BINOPS[type(node.op)](self._eval(node.left, lam), self._eval(node.right, lam))
        if isinstance(node, ast.UnaryOp):
            v = self._eval(node.operand, lam)
            return -v if isinstance(node.op, ast.USub) else v
        if isinstance(node, ast.Call):
            return self._funcs[node.func.id](*[self._eval(a, lam) for a in node.args])
        if isinstance(node, ast.Constant):
            return float(node.value)
        return lam

    def __call__(self, lam):
        return float(self._eval(self._tree, float(lam)))


def tabulate(text, n_lambda_steps, functions=None):
    """Values of the expression at lambda = k / n_lambda_steps, k = 0..n_lambda_steps."""
    e = Expression(text, functions)
    if n_lambda_steps <= 0:
        return [e(0.0)]
    return [e(k / float(n_lambda_steps)) for k in range(n_lambda_steps + 1)]


# ---------------------------------------------------------------------------------------------------------
# Compiler for Custom*Force energy expressions (OpenMM hands these to Lepton; blues/tests/data/ethylene_system.xml:52,
# 96 are the reference's use).  ``compile_program`` lowers ``"expr; name = expr; ..."`` to the stack program of
# include/blues_b200.h (BL_OP_*), which the engine evaluates together with its derivative in r (k_custom).
# ---------------------------------------------------------------------------------------------------------
OP = {'CONST': 0, 'R': 1, 'PARAM': 2, 'GLOBAL': 3, 'ADD': 4, 'SUB': 5, 'MUL': 6, 'DIV': 7, 'NEG': 8, 'POWI': 9, 'POW': 10,
      'SQRT': 11, 'EXP': 12, 'LOG': 13, 'SIN': 14, 'COS': 15, 'TAN': 16, 'ABS': 17, 'MIN': 18, 'MAX': 19, 'STEP': 20,
      'DELTA': 21, 'SELECT': 22, 'ERF': 23, 'ERFC': 24, 'TANH': 25, 'SINH': 26, 'COSH': 27, 'ATAN': 28}
_UNARY_FUNCS = {'sqrt': 'SQRT', 'exp': 'EXP', 'log': 'LOG', 'sin': 'SIN', 'cos': 'COS', 'tan': 'TAN', 'abs': 'ABS',
                'step': 'STEP', 'delta': 'DELTA', 'erf': 'ERF', 'erfc': 'ERFC', 'tanh': 'TANH', 'sinh': 'SINH',
                'cosh': 'COSH', 'atan': 'ATAN'}
_LAMBDA_GLOBALS = {'lambda_sterics': 0, 'lambda_electrostatics': 1}
MAX_STACK = 24


def _parse(src):
    src = re.sub(r'\blambda\b', _VAR, src.replace('^', '**'))
    return ast.parse(src.strip(), mode='eval').body


def compile_program(text, params=None, constants=None, distance_calls=()):
    """``text``: Lepton energy expression with optional ``; name = expr`` definitions.  ``params``: name → per-term
    parameter slot.  ``constants``: global parameters other than the two lambdas, folded in as immediates.
    ``distance_calls``: function-call spellings that stand for the term's distance (``distance(g1,g2)``); the bare name
    ``r`` always does.  Returns ``(ops, args)`` — two equally long lists."""
    params = dict(params or {})
    constants = dict(constants or {})
    parts = [p for p in str(text).split(';') if p.strip()]
    if not parts:
        raise ValueError('empty energy expression')
    defs = {}
    for p in parts[1:]:
        name, _, rhs = p.partition('=')
        if not _ or not name.strip().isidentifier():
            raise ValueError('malformed definition %r in %r' % (p, text))
        defs[name.strip()] = _parse(rhs)
    dist = set(re.sub(r'\s+', '', d) for d in distance_calls)
    ops, args = [], []
    depth = [0, 0]                                       # current, maximum

    def emit(op, arg=0.0, delta=1):
        ops.append(OP[op])
        args.append(float(arg))
        depth[0] += delta
        depth[1] = max(depth[1], depth[0])

    def const_value(node):
        if isinstance(node, ast.Constant) and isinstance(node.value, (int, float)):
            return float(node.value)
        if isinstance(node, ast.UnaryOp) and isinstance(node.op, (ast.USub, ast.UAdd)):
            v = const_value(node.operand)
            return None if v is None else (-v if isinstance(node.op, ast.USub) else v)
        return None

    active = []

    def gen(node):
        cv = const_value(node)
        if cv is not None:
            return emit('CONST', cv)
        if isinstance(node, ast.BinOp):
            if isinstance(node.op, ast.Pow):
                e = const_value(node.right)
                gen(node.left)
                if e is not None and e == int(e) and abs(e) <= 64:
                    return emit('POWI', int(e), 0)
                gen(node.right)
                return emit('POW', 0, -1)
            gen(node.left)
            gen(node.right)
            name = {ast.Add: 'ADD', ast.Sub: 'SUB', ast.Mult: 'MUL', ast.Div: 'DIV'}.get(type(node.op))
            if name is None:
                raise ValueError('unsupported operator in %r' % text)
            return emit(name, 0, -1)
        if isinstance(node, ast.UnaryOp) and isinstance(node.op, (ast.USub, ast.UAdd)):
            gen(node.operand)
            return emit('NEG', 0, 0) if isinstance(node.op, ast.USub) else None
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Name) and not node.keywords:
            f = node.func.id
            spelled = re.sub(r'\s+', '', ast.unparse(node))
            if spelled in dist:
                return emit('R')
            if f in _UNARY_FUNCS and len(node.args) == 1:
                gen(node.args[0])
                return emit(_UNARY_FUNCS[f], 0, 0)
            if f in ('min', 'max') and len(node.args) == 2:
                gen(node.args[0]); gen(node.args[1])
                return emit(f.upper(), 0, -1)
            if f == 'select' and len(node.args) == 3:
                for a in node.args:
                    gen(a)
                return emit('SELECT', 0, -2)
            if f in ('square', 'cube', 'recip') and len(node.args) == 1:
                gen(node.args[0])
                return emit('POWI', {'square': 2, 'cube': 3, 'recip': -1}[f], 0)
            raise NotImplementedError('function %s(...) of %r is not supported by the custom-force evaluator' % (f, text))
        if isinstance(node, ast.Name):
            n = node.id
            if n in defs:
                if n in active:
                    raise ValueError('circular definition of %s in %r' % (n, text))
                active.append(n)
                gen(defs[n])
                active.pop()
                return None
            if n == 'r':
                return emit('R')
            if n in params:
                return emit('PARAM', params[n])
            if n in _LAMBDA_GLOBALS:
                return emit('GLOBAL', _LAMBDA_GLOBALS[n])
            if n in constants:
                return emit('CONST', constants[n])
            raise ValueError('unknown name %r in energy expression %r' % (n, text))
        raise ValueError('unsupported element in energy expression %r' % text)

    gen(_parse(parts[0]))
    if depth[0] != 1 or depth[1] > MAX_STACK:
        raise ValueError('energy expression %r is too deep for the evaluator (%d > %d)' % (text, depth[1], MAX_STACK))
    return ops, args


def evaluate_program(ops, args, r, par=(), g=(1.0, 1.0)):
    """Host twin of the device interpreter (value and d/dr by dual numbers) — used by the CPU tests of the compiler."""
    inv = {v: k for k, v in OP.items()}
    st = []
    for op, a in zip(ops, args):
        k = inv[op]
        if k == 'CONST': st.append((a, 0.0))
        elif k == 'R': st.append((r, 1.0))
        elif k == 'PARAM': st.append((float(par[int(a)]), 0.0))
        elif k == 'GLOBAL': st.append((float(g[int(a)]), 0.0))
        elif k in ('ADD', 'SUB', 'MUL', 'DIV', 'POW', 'MIN', 'MAX'):
            (y, dy), (x, dx) = st.pop(), st.pop()
            if k == 'ADD': st.append((x + y, dx + dy))
            elif k == 'SUB': st.append((x - y, dx - dy))
            elif k == 'MUL': st.append((x * y, dx * y + x * dy))
            elif k == 'DIV': st.append((x / y, (dx - x / y * dy) / y))
            elif k == 'POW':
                v = x ** y
                st.append((v, (y * x ** (y - 1) * dx if dx else 0.0) + (v * math.log(x) * dy if dy else 0.0)))
            elif k == 'MIN': st.append((y, dy) if y < x else (x, dx))
            else: st.append((y, dy) if y > x else (x, dx))
        elif k == 'SELECT':
            (b, db), (a_, da), (c, _) = st.pop(), st.pop(), st.pop()
            st.append((a_, da) if c != 0 else (b, db))
        else:
            x, dx = st.pop()
            if k == 'NEG': st.append((-x, -dx))
            elif k == 'POWI':
                n = int(a)
                st.append((1.0, 0.0) if n == 0 else (x ** n, n * x ** (n - 1) * dx))
            elif k == 'SQRT': st.append((math.sqrt(x), 0.5 * dx / math.sqrt(x) if dx else 0.0))
            elif k == 'EXP': st.append((math.exp(x), math.exp(x) * dx))
            elif k == 'LOG': st.append((math.log(x), dx / x))
            elif k == 'SIN': st.append((math.sin(x), math.cos(x) * dx))
            elif k == 'COS': st.append((math.cos(x), -math.sin(x) * dx))
            elif k == 'TAN': st.append((math.tan(x), (1 + math.tan(x) ** 2) * dx))
            elif k == 'ABS': st.append((abs(x), dx if x >= 0 else -dx))
            elif k == 'STEP': st.append((1.0 if x >= 0 else 0.0, 0.0))
            elif k == 'DELTA': st.append((1.0 if x == 0 else 0.0, 0.0))
            elif k == 'ERF': st.append((math.erf(x), 2 / math.sqrt(math.pi) * math.exp(-x * x) * dx))
            elif k == 'ERFC': st.append((math.erfc(x), -2 / math.sqrt(math.pi) * math.exp(-x * x) * dx))
            elif k == 'TANH': st.append((math.tanh(x), (1 - math.tanh(x) ** 2) * dx))
            elif k == 'SINH': st.append((math.sinh(x), math.cosh(x) * dx))
            elif k == 'COSH': st.append((math.cosh(x), math.sinh(x) * dx))
            elif k == 'ATAN': st.append((math.atan(x), dx / (1 + x * x)))
    return st[0]
