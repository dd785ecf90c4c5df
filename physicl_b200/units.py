"""Code-units system: a ``numpy.ndarray`` subclass that carries physical units.

This is the host-side mirror of the reference's ``Measurement`` (physicl/__init__.py:18-291) so
that user code written against PhysiCL keeps working; it is off the hot path.  The kernels only ever
see plain floats: host steps call ``float(x)`` / ``x.__unscaled__()`` at the C-ABI boundary.

Behaviour kept from the reference (its test/test_units.py is the specification):
  * the stored numbers are in CODE units: ``raw * unit_factor * prod(code_scale[base]**power)``
  * ``units`` maps code dimensions (L, T, M, I, Th, N, J) to powers; ``original_units`` remembers
    the spelling the user gave, and plain numbers mixed into arithmetic are read in those units
  * add/subtract keep the left operand's units, multiply/divide combine them, power/sqrt scale them
  * ``str(x)`` is the upper-cased ndarray repr, which is what the reference splices into kernels
"""
from __future__ import annotations

import copy
import operator
import re

import numpy as np


class MeasurementError(ArithmeticError):
    pass


# base unit -> [code scale, (code dimension, power)]
_BASE = ("s", "T"), ("m", "L"), ("kg", "M"), ("A", "I"), ("K", "Th"), ("mol", "N"), ("cd", "J")

# derived / accepted units -> (factor, ((unit, power), ...)), all expressed through other table entries
_DERIVED = {
    "N": (1, (("kg", 1), ("m", 1), ("s", -2))),
    "Pa": (1, (("kg", 1), ("m", -1), ("s", -2))),
    "J": (1, (("N", 1), ("m", 1))),
    "W": (1, (("kg", 1), ("m", 2), ("s", -3))),
    "C": (1, (("A", 1), ("s", 1))),
    "V": (1, (("W", 1), ("A", -1))),
    "F": (1, (("C", 1), ("V", -1))),
    "Ohm": (1, (("V", 1), ("A", 1))),
    "Wb": (1, (("V", 1), ("s", 1))),
    "T": (1, (("Wb", 1), ("m", -2))),
    "H": (1, (("Wb", 1), ("A", -1))),
    "lm": (1, (("cd", 1),)),
    "Bq": (1, (("s", -1),)),
    "Gy": (1, (("m", 2), ("s", -2))),
    "Sv": (1, (("m", 2), ("s", -2))),
    "kat": (1, (("mol", 1), ("s", -1))),
    "min": (60, (("s", 1),)),
    "h": (3600, (("s", 1),)),
    "d": (86400, (("s", 1),)),
    "au": (149597870700, (("m", 1),)),
    "ha": (10 ** 4, (("m", 2),)),
    "L": (10 ** -3, (("m", 3),)),
    "t": (10 ** 3, (("kg", 1),)),
    "Da": (1.6605390666050e-27, (("kg", 1),)),
    "eV": (1.602176634e-19, (("J", 1),)),
}

_PARSED = {}  # Measurement._apply_units: parsed unit strings
_ONE = np.double(1)
_FIRST = operator.itemgetter(0)
_TOKEN = re.compile(r"([a-zA-Z]+)\s*(?:\*\*|\^)\s*(-?\d+(?:\.\d+)?)")
_DIVIDES = ("divide", "true_divide", "floor_divide")


class Measurement(np.ndarray):
    # public tables with the reference's shapes: name -> [scale, (unit, power), ...]
    code_scale = {u: [1, (dim, 1)] for u, dim in _BASE}
    unit_scale = {**{u: [1, (u, 1)] for u, _ in _BASE}, **{k: [f, *parts] for k, (f, parts) in _DERIVED.items()}}

    # ---- unit algebra -----------------------------------------------------------------------
    @staticmethod
    def _to_base(unit, power):
        """unit**power -> (numeric factor, [(base unit, power), ...])"""
        if unit not in Measurement.unit_scale:
            raise MeasurementError("unknown unit '%s'" % unit)
        entry = Measurement.unit_scale[unit]
        factor = entry[0] ** power
        out = []
        for sub, p in entry[1:]:
            if sub in Measurement.code_scale:
                out.append((sub, p * power))
            else:
                # the reference keeps only the top-level factor of nested derived units
                # (physicl/__init__.py:107-112); every nested factor in its table is 1 except eV -> J
                out.extend(Measurement._to_base(sub, p * power)[1])
        return factor, out

    @staticmethod
    def set_code_scale(base_unit, new_scale):
        """Choose how many code units one SI base unit is worth (physicl/__init__.py:125-126)."""
        Measurement.code_scale[base_unit][0] = new_scale

    @staticmethod
    def reset_code_scale(base_unit):
        Measurement.set_code_scale(base_unit, 1)

    def __new__(cls, raw_value, units):
        if isinstance(raw_value, list):
            raw_value = [v.__unscaled__() if isinstance(v, Measurement) else v for v in raw_value]
        elif isinstance(raw_value, Measurement):
            raw_value = raw_value.view(np.ndarray)
        obj = np.array(raw_value, dtype=np.double).view(cls)
        obj._apply_units(units)
        return obj

    @staticmethod
    def _parse_units(spec):
        """unit string -> (scale, dimensions, units as spelled) under the current code scales"""
        scale = np.double(1)
        dims, spelled = {}, {}
        for unit, power in _TOKEN.findall(spec or ""):
            power = float(power)
            power = int(power) if power == int(power) else power
            factor, base = Measurement._to_base(unit, power)
            scale = scale * factor
            for b, p in base:
                cs, (dim, dp) = Measurement.code_scale[b]
                scale = scale * cs ** p
                dims[dim] = dims.get(dim, 0) + dp * p
            spelled[unit] = spelled.get(unit, 0) + power
        return scale, dims, spelled

    def _apply_units(self, spec):
        # Object.__init__ parses the same five strings for every particle: remember the result per (string, code
        # scales, unit table size); set_code_scale and new entries in unit_scale change the key
        key = (spec, tuple(map(_FIRST, Measurement.code_scale.values())), len(Measurement.unit_scale))
        hit = _PARSED.get(key)
        if hit is None:
            if len(_PARSED) > 4096:
                _PARSED.clear()
            hit = _PARSED[key] = Measurement._parse_units(spec)
        scale, dims, spelled = hit
        self.scale, self.units, self.original_units = scale, dict(dims), dict(spelled)
        if scale != 1:
            raw = self.view(np.ndarray)
            np.multiply(raw, scale, out=raw)

    __scale__ = _apply_units  # the reference's name for it (physicl/__init__.py:141)

    def rescale(self):
        """physicl/__init__.py:289-291: a placeholder there too (one global code scale, no per-value rescaling)."""

    def __array_finalize__(self, src):
        if src is None:
            return
        if type(src) is np.ndarray:  # a plain array viewed as a Measurement: dimensionless until told otherwise
            self.scale, self.units, self.original_units = _ONE, {}, {}
            return
        self.scale = getattr(src, "scale", np.double(1))
        self.units = dict(getattr(src, "units", {}) or {})
        self.original_units = dict(getattr(src, "original_units", {}) or {})

    # ---- views ------------------------------------------------------------------------------
    def __unscaled__(self):
        return np.array(self.view(np.ndarray), dtype=np.double) / self.scale

    def value(self):
        return self.__unscaled__()

    def unitstr(self):
        try:
            return " ".join("%s**%s" % (k, int(v) if v == int(v) else float(v)) for k, v in self.original_units.items())
        except AttributeError:
            return ""

    def fstr(self):
        return str(float(self))

    def valstr(self):
        return str(self.value())

    def __str__(self):
        return str(self.view(np.ndarray)).upper()

    def __format__(self, fmt):
        return self.view(np.ndarray).__format__(fmt).upper()

    def __repr__(self):
        return str(self.value()) + " " + self.unitstr()

    def __deepcopy__(self, memo):
        out = np.array(self.view(np.ndarray), dtype=np.double).view(Measurement)
        out.scale = self.scale
        out.units = copy.deepcopy(self.units, memo)
        out.original_units = copy.deepcopy(self.original_units, memo)
        return out

    def __reduce__(self):
        fn, args, state = super().__reduce__()
        return fn, args, (state, self.scale, self.units, self.original_units)

    def __setstate__(self, state):
        base, self.scale, self.units, self.original_units = state
        super().__setstate__(base)

    # ---- arithmetic -------------------------------------------------------------------------
    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        lead = inputs[0] if isinstance(inputs[0], Measurement) else next(i for i in inputs if isinstance(i, Measurement))
        # plain numbers are read in the leading operand's (user-spelled) units
        conv = [i if isinstance(i, Measurement) and hasattr(i, "units") else Measurement(i, lead.unitstr()) for i in inputs]
        raw = [c.view(np.ndarray) for c in conv]
        if "out" in kwargs:
            kwargs["out"] = tuple(o.view(np.ndarray) if isinstance(o, np.ndarray) else o for o in kwargs["out"])
        val = getattr(ufunc, method)(*raw, **kwargs)
        name = ufunc.__name__
        if isinstance(val, tuple) or val is None or method != "__call__":
            return val
        first = conv[0]
        if name in ("multiply",) + _DIVIDES and len(conv) == 2:
            sgn = -1 if name in _DIVIDES else 1
            other = conv[1]
            dims = dict(first.units)
            spelled = dict(first.original_units)
            for k, p in other.units.items():
                dims[k] = dims.get(k, 0) + sgn * p
            for k, p in other.original_units.items():
                spelled[k] = spelled.get(k, 0) + sgn * p
            out = np.asarray(val, dtype=np.double).view(Measurement)
            out.scale, out.units, out.original_units = first.scale * other.scale ** sgn, dims, spelled
            return out
        out = np.asarray(val).view(Measurement)
        if name in ("power", "square", "sqrt"):
            p = {"square": 2, "sqrt": 0.5}.get(name)
            if p is None:
                p = raw[1]
                p = p.item() if np.ndim(p) == 0 else p
            out.scale = first.scale ** p
            out.units = {k: v * p for k, v in first.units.items()}
            out.original_units = {k: v * p for k, v in first.original_units.items()}
        else:
            out.scale = first.scale
            out.units = copy.deepcopy(first.units)
            out.original_units = copy.deepcopy(first.original_units)
        return out
