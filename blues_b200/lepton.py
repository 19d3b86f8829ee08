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


class Expression(object):
    def __init__(self, text):
        self.text = str(text)
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
        elif isinstance(node, ast.Call) and isinstance(node.func, ast.Name) and node.func.id in _FUNCS and not node.keywords:
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
            return _FUNCS[node.func.id](*[self._eval(a, lam) for a in node.args])
        if isinstance(node, ast.Constant):
            return float(node.value)
        return lam

    def __call__(self, lam):
        return float(self._eval(self._tree, float(lam)))


def tabulate(text, n_lambda_steps):
    """Values of the expression at lambda = k / n_lambda_steps, k = 0..n_lambda_steps."""
    e = Expression(text)
    if n_lambda_steps <= 0:
        return [e(0.0)]
    return [e(k / float(n_lambda_steps)) for k in range(n_lambda_steps + 1)]
