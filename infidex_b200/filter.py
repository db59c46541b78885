"""Infiscript filters on the host: a compact AST -> stack bytecode compiler and the INFISCRIPT-V1 serializer.

Mirrors /root/reference/src/Infidex:
  Api/FilterParser.cs:444-449      comparison -> ValueFilter / RangeFilter / NOT(ValueFilter)
  Filtering/FilterCompiler.cs:84-279  AND/OR short-circuit layout (DUP, JUMP_IF_*, POP), BETWEEN, IN, string ops, NULL checks
  Filtering/ConstantPool.cs:75-112, BytecodeSerializer.cs:16-62   wire format consumed by ifx_filter_register
The text parser here covers the expression subset used by the configs (comparisons, AND/OR/NOT, parentheses,
BETWEEN, IN, CONTAINS/STARTS WITH/ENDS WITH/LIKE, IS [NOT] NULL); the full Infiscript grammar (ternaries, regex)
stays with the C# host, which hands over bytecode.
"""
import re
import struct

OP = {"PUSH_FIELD": 0x01, "PUSH_CONST": 0x02, "POP": 0x03, "DUP": 0x04, "EQ": 0x10, "NEQ": 0x11, "LT": 0x12, "LTE": 0x13, "GT": 0x14,
      "GTE": 0x15, "AND": 0x20, "OR": 0x21, "NOT": 0x22, "CONTAINS": 0x30, "STARTS_WITH": 0x31, "ENDS_WITH": 0x32, "LIKE": 0x33,
      "MATCHES": 0x34, "IN": 0x40, "BETWEEN": 0x41, "IS_NULL": 0x50, "IS_NOT_NULL": 0x51, "JUMP": 0x60, "JUMP_IF_FALSE": 0x61,
      "JUMP_IF_TRUE": 0x62, "HALT": 0xFF}
_HAS_OPERAND = {0x01, 0x02, 0x60, 0x61, 0x62}


class _Pool:
    def __init__(self):
        self.items, self.index = [], {}

    def add_string(self, s):
        k = ("s", s)
        if k not in self.index:
            self.index[k] = len(self.items); self.items.append(("s", s))
        return self.index[k]

    def add_number(self, x):
        k = ("n", float(x))
        if k not in self.index:
            self.index[k] = len(self.items); self.items.append(("n", float(x)))
        return self.index[k]

    def add_array(self, values):
        self.items.append(("a", [str(v) for v in values])); return len(self.items) - 1

    @staticmethod
    def _dotnet_string(s):   # BinaryWriter.Write(string): 7-bit encoded byte length + UTF-8
        b = s.encode("utf-8"); n = len(b); out = bytearray()
        while n >= 0x80:
            out.append((n & 0x7F) | 0x80); n >>= 7
        out.append(n); return bytes(out) + b

    def serialize(self):
        out = bytearray(struct.pack("<i", len(self.items)))
        for kind, v in self.items:
            if kind == "s":
                out += b"\x01" + self._dotnet_string(v)
            elif kind == "n":
                out += b"\x02" + struct.pack("<d", v)
            else:
                out += b"\x03" + struct.pack("<i", len(v)) + b"".join(self._dotnet_string(x) for x in v)
        return bytes(out)


class FilterParseError(ValueError):
    """Api/FilterParseException.cs"""


class Filter:
    """AST node. kind in: value, range, in, string, null, and, or, not, ternary, literal."""

    def __init__(self, kind, **kw):
        self.kind = kind; self.__dict__.update(kw); self._code = None

    # ---- constructors mirroring the reference's filter classes
    @staticmethod
    def Value(field, value): return Filter("value", field=field, value=value)

    @staticmethod
    def Range(field, min=None, max=None, include_min=True, include_max=True):
        return Filter("range", field=field, min=min, max=max, include_min=include_min, include_max=include_max)

    @staticmethod
    def In(field, values): return Filter("in", field=field, values=list(values))

    @staticmethod
    def String(field, op, pattern): return Filter("string", field=field, op=op, pattern=pattern)

    @staticmethod
    def Null(field, is_null=True): return Filter("null", field=field, is_null=is_null)

    @staticmethod
    def And(a, b): return Filter("and", left=a, right=b)

    @staticmethod
    def Or(a, b): return Filter("or", left=a, right=b)

    @staticmethod
    def Not(a): return Filter("not", left=a)

    @staticmethod
    def Ternary(cond, true_value, false_value): return Filter("ternary", cond=cond, left=true_value, right=false_value)

    @staticmethod
    def Literal(value): return Filter("literal", value=value)

    def fields(self):
        """Names of the document fields the filter reads."""
        out = set()
        if self.kind in ("and", "or"): out |= self.left.fields() | self.right.fields()
        elif self.kind == "not": out |= self.left.fields()
        elif self.kind == "ternary": out |= self.cond.fields() | self.left.fields() | self.right.fields()
        elif self.kind == "literal": pass
        else: out.add(self.field)
        return out

    # ---- FilterCompiler
    def _compile(self, pool, code):
        k = self.kind
        if k == "and" or k == "or":
            self.left._compile(pool, code); code.append([OP["DUP"], None]); jp = len(code)
            code.append([OP["JUMP_IF_FALSE" if k == "and" else "JUMP_IF_TRUE"], 0]); code.append([OP["POP"], None])
            self.right._compile(pool, code); code[jp][1] = len(code)
        elif k == "not":
            self.left._compile(pool, code); code.append([OP["NOT"], None])
        elif k == "ternary":      # FilterCompiler.CompileTernary (:224-252)
            self.cond._compile(pool, code); jf = len(code); code.append([OP["JUMP_IF_FALSE"], 0]); code.append([OP["POP"], None])
            self.left._compile(pool, code); je = len(code); code.append([OP["JUMP"], 0])
            code[jf][1] = len(code); code.append([OP["POP"], None]); self.right._compile(pool, code); code[je][1] = len(code)
        elif k == "literal":      # CompileLiteral (:254-279)
            v = self.value
            if isinstance(v, bool): c = pool.add_string("True" if v else "False")
            elif isinstance(v, (int, float)): c = pool.add_number(v)
            else: c = pool.add_string("null" if v is None else str(v))
            code.append([OP["PUSH_CONST"], c])
        elif k == "value":
            f = pool.add_string(self.field); v = pool.add_string("" if self.value is None else str(self.value))
            code += [[OP["PUSH_FIELD"], f], [OP["PUSH_CONST"], v], [OP["EQ"], None]]
        elif k == "range":
            f = pool.add_string(self.field)
            if self.min is not None and self.max is not None:
                a = pool.add_string(str(self.min)); b = pool.add_string(str(self.max))
                code += [[OP["PUSH_FIELD"], f], [OP["PUSH_CONST"], a], [OP["PUSH_CONST"], b], [OP["BETWEEN"], None]]
            elif self.min is not None:
                a = pool.add_string(str(self.min)); code += [[OP["PUSH_FIELD"], f], [OP["PUSH_CONST"], a], [OP["GTE" if self.include_min else "GT"], None]]
            elif self.max is not None:
                b = pool.add_string(str(self.max)); code += [[OP["PUSH_FIELD"], f], [OP["PUSH_CONST"], b], [OP["LTE" if self.include_max else "LT"], None]]
        elif k == "in":
            f = pool.add_string(self.field); a = pool.add_array(self.values)
            code += [[OP["PUSH_FIELD"], f], [OP["PUSH_CONST"], a], [OP["IN"], None]]
        elif k == "string":
            f = pool.add_string(self.field); p = pool.add_string(self.pattern)
            code += [[OP["PUSH_FIELD"], f], [OP["PUSH_CONST"], p], [OP[self.op], None]]
        elif k == "null":
            f = pool.add_string(self.field); code += [[OP["PUSH_FIELD"], f], [OP["IS_NULL" if self.is_null else "IS_NOT_NULL"], None]]
        else:
            raise ValueError(k)

    def bytecode(self):
        if self._code is None:
            pool, code = _Pool(), []
            self._compile(pool, code); code.append([OP["HALT"], None])
            pb = pool.serialize()
            out = bytearray(b"INFISCRIPT-V1" + struct.pack("<H", 1) + struct.pack("<i", len(pb)) + pb + struct.pack("<i", len(code)))
            for op, a in code:
                out.append(op)
                if op in _HAS_OPERAND:
                    out += struct.pack("<i", a)
            self._code = bytes(out)
        return self._code

    def __hash__(self): return hash(self.bytecode())

    def __eq__(self, o): return isinstance(o, Filter) and self.bytecode() == o.bytecode()

    # ---- Filter.Parse: the Infiscript grammar (Api/Infiscript.bnf, Api/FilterParser.cs)
    _TOK = re.compile(r"\s*(?:(<=|>=|!=|&&|\|\||=|<|>|\(|\)|,|&|\||!|\?|:)|'((?:[^']|'')*)'|\"((?:[^\"]|\"\")*)\"|([A-Za-z_][A-Za-z0-9_\.]*)|(-?\d+(?:\.\d+)?(?:[eE][+-]?\d+)?))")

    @staticmethod
    def Parse(expr):
        toks, pos = [], 0
        while pos < len(expr):
            if expr[pos:].strip() == "":
                break
            m = Filter._TOK.match(expr, pos)
            if not m:
                raise FilterParseError("cannot tokenize filter at %d: %r" % (pos, expr[pos:pos + 20]))
            pos = m.end()
            if m.group(1): toks.append(("op", m.group(1)))
            elif m.group(2) is not None: toks.append(("val", m.group(2).replace("''", "'")))
            elif m.group(3) is not None: toks.append(("val", m.group(3).replace('""', '"')))
            elif m.group(4): toks.append(("id", m.group(4)))
            else: toks.append(("val", m.group(5)))
        p = [0]

        def peek(): return toks[p[0]] if p[0] < len(toks) else (None, None)

        def kw(word): t = peek(); return t[0] == "id" and t[1].upper() == word

        def take(): t = peek(); p[0] += 1; return t

        def is_op(*ops): return peek()[0] == "op" and peek()[1] in ops

        def need(cond, msg):
            if not cond:
                raise FilterParseError(msg)

        def value():
            t = take(); need(t[0] in ("val", "id"), "expected a value"); return t[1]

        def parse_ternary():        # <or_expression> [ "?" <ternary> ":" <ternary> ], right-associative, lowest precedence
            cond = parse_or()
            if is_op("?"):
                take(); tv = parse_ternary(); need(is_op(":"), "ternary format is: condition ? true_value : false_value"); take(); fv = parse_ternary()
                return Filter.Ternary(cond, tv, fv)
            return cond

        def parse_or():
            left = parse_and()
            while kw("OR") or is_op("||", "|"):
                take(); left = Filter.Or(left, parse_and())
            return left

        def parse_and():
            left = parse_not()
            while kw("AND") or is_op("&&", "&"):
                take(); left = Filter.And(left, parse_not())
            return left

        def parse_not():
            if kw("NOT") or is_op("!"):
                take(); return Filter.Not(parse_atom())      # <not_operator> <primary_expression>
            return parse_atom()

        def parse_atom():
            t = take()
            if t == ("op", "("):
                e = parse_ternary(); need(take() == ("op", ")"), "expected )"); return e
            if t[0] == "val":            # literal (ternary branches): a number when it parses as one (FilterParser.cs:179-191)
                try: return Filter.Literal(float(t[1]))
                except ValueError: return Filter.Literal(t[1])
            need(t[0] == "id", "expected field name")
            field = t[1]
            if kw("BETWEEN"):
                take(); a = value(); need(kw("AND"), "expected AND in BETWEEN"); take(); b = value(); return Filter.Range(field, a, b)
            if kw("IN"):
                take(); need(take() == ("op", "("), "expected ( after IN"); vals = [value()]
                while is_op(","):
                    take(); vals.append(value())
                need(take() == ("op", ")"), "expected ) after IN list"); return Filter.In(field, vals)
            if kw("IS"):
                take(); neg = False
                if kw("NOT"): take(); neg = True
                need(kw("NULL"), "expected NULL"); take(); return Filter.Null(field, not neg)
            for word, op in (("CONTAINS", "CONTAINS"), ("LIKE", "LIKE"), ("MATCHES", "MATCHES")):
                if kw(word):
                    take(); return Filter.String(field, op, value())
            if kw("STARTS"):
                take(); need(kw("WITH"), "expected WITH"); take(); return Filter.String(field, "STARTS_WITH", value())
            if kw("ENDS"):
                take(); need(kw("WITH"), "expected WITH"); take(); return Filter.String(field, "ENDS_WITH", value())
            op = take(); need(op[0] == "op" and op[1] in ("=", "!=", ">", ">=", "<", "<="), "expected comparison operator"); v = value()
            return {"=": lambda: Filter.Value(field, v), "!=": lambda: Filter.Not(Filter.Value(field, v)),
                    ">": lambda: Filter.Range(field, min=v, include_min=False), ">=": lambda: Filter.Range(field, min=v, include_min=True),
                    "<": lambda: Filter.Range(field, max=v, include_max=False), "<=": lambda: Filter.Range(field, max=v, include_max=True)}[op[1]]()

        if not toks:
            raise FilterParseError("empty filter expression")
        e = parse_ternary()
        if p[0] != len(toks):
            raise FilterParseError("trailing tokens in filter expression")
        return e
