"""Equilibrium equations -> device bytecode (the host half of the reference's compile_residuals, ModelParser.jl:217-259).

The reference turns each YAML equation string "LHS = RHS" into Julia code that evaluates LHS .- RHS over the columns of
the padded variable matrix xMat (rows = var_names(model): endogenous, heterogeneous, exogenous; `VAR(-k)` / `VAR(+k)`
read the row shifted by k periods, with the steady-state boundary columns of assemble_full_xMat,
GeneralStructures.jl:329-377, beyond the ends).  Here the same strings are compiled to a small postfix program that the
device interprets once per (period, equation) — and once per (lane, period, equation) with dual arithmetic for the
JVPs — so the aggregate block of a model is data, not a built-in kernel (csrc/hank_eq.cu, hank_eq_configure).

Grammar: + - * / ^, unary minus, parentheses, numbers, parameter names, variable names with an optional integer shift
`VAR(-1)`, and the functions exp, log, sqrt.  n-ary products / sums fold to the left like transform_expr does
(ModelParser.jl:80-110), `^` is right-associative and binds tighter than unary minus, as in Julia.
"""
import ast
import re

import numpy as np

# opcodes (csrc/hank_eq.cu keeps the same numbering)
OP_CONST, OP_VAR, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_POW, OP_NEG, OP_EXP, OP_LOG, OP_SQRT = range(11)
_BIN = {ast.Add: OP_ADD, ast.Sub: OP_SUB, ast.Mult: OP_MUL, ast.Div: OP_DIV, ast.Pow: OP_POW}
_FUN = {"exp": OP_EXP, "log": OP_LOG, "sqrt": OP_SQRT}
MAX_STACK = 16


class EquationProgram:
    """Postfix programs of n_eq equations over the variables `names` (endogenous first, then the heterogeneous
    aggregates, then the exogenous variables — var_names(model), ModelParser.jl:357)."""

    def __init__(self, equations, names, params):
        self.names = tuple(names)
        self.equations = tuple(equations)
        self.consts = []
        self._cidx = {}
        self.code = []
        self.eq_off = [0]
        self.max_lag = 0
        self.max_lead = 0
        var_idx = {n: i for i, n in enumerate(self.names)}
        for eq in self.equations:
            parts = eq.split("=")
            if len(parts) != 2:
                raise ValueError(f"Equation must contain exactly one '=': {eq}")
            depth = self._emit(self._parse(parts[0], var_idx), var_idx, params, 0)
            depth = max(depth, 1 + self._emit(self._parse(parts[1], var_idx), var_idx, params, 0))
            self.code.append(OP_SUB)
            if depth > MAX_STACK:
                raise ValueError(f"equation needs an evaluation stack deeper than {MAX_STACK}: {eq}")
            self.eq_off.append(len(self.code))

    # VAR(-1) is a call in Python's grammar too; shifts are rewritten so that '+1' parses
    def _parse(self, text, var_idx):
        text = text.strip().replace("^", "**")
        names = "|".join(sorted((re.escape(n) for n in var_idx), key=len, reverse=True))
        text = re.sub(rf"\b({names})\s*\(\s*([+-]?\d+)\s*\)", lambda m: f"__shift__({m.group(1)}, {int(m.group(2))})", text)
        return ast.parse(text, mode="eval").body

    def _const(self, v):
        v = float(v)
        key = np.float64(v).tobytes()
        if key not in self._cidx:
            self._cidx[key] = len(self.consts)
            self.consts.append(v)
        return self._cidx[key]

    def _emit(self, node, var_idx, params, depth):
        """Appends the node's postfix code; returns the deepest stack it reaches (its result occupies slot `depth`)."""
        if isinstance(node, ast.Constant) and isinstance(node.value, (int, float)):
            self.code += [OP_CONST, self._const(node.value)]
            return depth + 1
        if isinstance(node, ast.Name):
            if node.id in var_idx:
                self.code += [OP_VAR, var_idx[node.id], 0]
            elif node.id in params:
                self.code += [OP_CONST, self._const(params[node.id])]
            else:
                raise ValueError(f"unknown symbol {node.id!r} (neither a variable nor a parameter)")
            return depth + 1
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Name):
            if node.func.id == "__shift__":
                v, k = node.args[0].id, int(ast.literal_eval(node.args[1]))
                self.max_lag = max(self.max_lag, -k); self.max_lead = max(self.max_lead, k)
                self.code += [OP_VAR, var_idx[v], k]
                return depth + 1
            if node.func.id in _FUN and len(node.args) == 1:
                d = self._emit(node.args[0], var_idx, params, depth)
                self.code.append(_FUN[node.func.id])
                return d
            raise ValueError(f"function {node.func.id!r} is not supported (exp, log, sqrt)")
        if isinstance(node, ast.UnaryOp) and isinstance(node.op, (ast.USub, ast.UAdd)):
            d = self._emit(node.operand, var_idx, params, depth)
            if isinstance(node.op, ast.USub):
                self.code.append(OP_NEG)
            return d
        if isinstance(node, ast.BinOp) and type(node.op) in _BIN:
            d1 = self._emit(node.left, var_idx, params, depth)
            d2 = self._emit(node.right, var_idx, params, depth + 1)
            self.code.append(_BIN[type(node.op)])
            return max(d1, d2)
        raise ValueError(f"unsupported expression: {ast.dump(node)}")

    @property
    def n_eq(self):
        return len(self.equations)

    def arrays(self):
        return (np.asarray(self.eq_off, dtype=np.int32), np.asarray(self.code, dtype=np.int32),
                np.asarray(self.consts if self.consts else [0.0], dtype=np.float64))
