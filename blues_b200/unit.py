"""Minimal unit/Quantity algebra with the subset of the ``simtk.unit`` surface that BLUES touches.

The reference passes ``simtk.unit.Quantity`` objects everywhere (``blues/simulation.py:432,968``,
``blues/moves.py:292-309``, ``blues/settings.py:140-187``).  This module provides the same surface
(``_value``, ``.unit``, ``value_in_unit``, ``*``/``/`` with units, numpy-array payloads with fancy
indexing) so user scripts and ``Move`` subclasses written for the reference keep working.  It is an
independent implementation: a unit is a (scale-to-SI, dimension-exponent) pair.
"""
import math
import numpy as _np

_DIMS = ('mass', 'length', 'time', 'temperature', 'charge', 'amount', 'angle')


class Unit(object):
    __array_priority__ = 200

    def __init__(self, factor, dims, name=None):
        self.factor = float(factor)
        self.dims = tuple(dims)
        self._name = name

    # -- algebra ---------------------------------------------------------------------------
    def __mul__(self, other):
        if isinstance(other, Unit):
            return Unit(self.factor * other.factor, [a + b for a, b in zip(self.dims, other.dims)],
                        _join(self._name, other._name, '*'))
        if isinstance(other, Quantity):
            return Quantity(other._value, other.unit * self)
        return Quantity(other, self)

    __rmul__ = lambda self, other: Quantity(other, self) if not isinstance(other, (Unit, Quantity)) else other.__mul__(self)

    def __truediv__(self, other):
        if isinstance(other, Unit):
            return Unit(self.factor / other.factor, [a - b for a, b in zip(self.dims, other.dims)],
                        _join(self._name, other._name, '/'))
        if isinstance(other, Quantity):
            return Quantity(1.0 / other._value, self / other.unit)
        return Quantity(1.0 / other, self)

    def __rtruediv__(self, other):
        inv = Unit(1.0 / self.factor, [-a for a in self.dims], _join('1', self._name, '/'))
        if isinstance(other, Quantity):
            return Quantity(other._value, other.unit * inv)
        return Quantity(other, inv)

    __div__ = __truediv__
    __rdiv__ = __rtruediv__

    def __pow__(self, p):
        return Unit(self.factor ** p, [a * p for a in self.dims], '(%s)**%s' % (self._name, p))

    def sqrt(self):
        return self ** 0.5

    def is_compatible(self, other):
        return all(abs(a - b) < 1e-12 for a, b in zip(self.dims, other.dims))

    def is_dimensionless(self):
        return all(abs(a) < 1e-12 for a in self.dims)

    def conversion_factor_to(self, other):
        if not self.is_compatible(other):
            raise TypeError('Unit "%s" is not compatible with Unit "%s".' % (self, other))
        return float('%.15g' % (self.factor / other.factor))

    def __eq__(self, other):
        return isinstance(other, Unit) and self.is_compatible(other) and math.isclose(
            self.factor, other.factor, rel_tol=1e-12)

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash((round(math.log(self.factor), 9), self.dims))

    def get_name(self):
        return self._name or 'dimensionless'

    def __str__(self):
        return self.get_name()

    __repr__ = __str__


def _join(a, b, op):
    a = a or '1'
    b = b or '1'
    if op == '*':
        return '%s*%s' % (a, b)
    return '%s/%s' % (a, b if ('*' not in b and '/' not in b) else '(%s)' % b)


def _base(name, factor, **exps):
    return Unit(factor, [exps.get(d, 0) for d in _DIMS], name)


dimensionless = _base(None, 1.0)


class Quantity(object):
    """A value (scalar, list or numpy array) tagged with a :class:`Unit`."""
    __array_priority__ = 100

    def __init__(self, value=None, unit=None):
        if isinstance(unit, Quantity):           # e.g. Quantity(3, 1/picoseconds)
            value = value * unit._value
            unit = unit.unit
        if unit is None:
            if isinstance(value, Quantity):
                value, unit = value._value, value.unit
            elif isinstance(value, (list, tuple)) and len(value) and isinstance(value[0], Quantity):
                unit = value[0].unit
                value = [v.value_in_unit(unit) for v in value]
            else:
                unit = dimensionless
        elif isinstance(value, Quantity):
            unit = value.unit * unit
            value = value._value
        elif isinstance(value, (list, tuple)) and len(value) and isinstance(value[0], Quantity):
            u0 = value[0].unit
            value = [v.value_in_unit(u0) for v in value]
            unit = u0 * unit
        if isinstance(value, (list, tuple)):
            if len(value) and isinstance(value[0], (list, tuple, _np.ndarray)):
                value = _np.array(value, dtype=float)
        self._value = value
        self.unit = unit

    # -- conversion ------------------------------------------------------------------------
    def value_in_unit(self, unit):
        f = self.unit.conversion_factor_to(unit)
        v = self._value
        if f == 1.0:
            return v
        if isinstance(v, (list, tuple)):
            return type(v)(x * f for x in v)
        return v * f

    def in_units_of(self, unit):
        return Quantity(self.value_in_unit(unit), unit)

    def value_in_unit_system(self, system=None):
        return self.value_in_unit(_md_unit_for(self.unit))

    def _reduce(self, value, unit):
        if unit.is_dimensionless():
            return value * unit.factor
        return Quantity(value, unit)

    # -- arithmetic ------------------------------------------------------------------------
    def __mul__(self, other):
        if isinstance(other, Unit):
            return self._reduce(self._value, self.unit * other)
        if isinstance(other, Quantity):
            return self._reduce(_arr(self._value) * _arr(other._value), self.unit * other.unit)
        return Quantity(_arr(self._value) * other, self.unit)

    def __rmul__(self, other):
        return Quantity(other * _arr(self._value), self.unit)

    def __truediv__(self, other):
        if isinstance(other, Unit):
            return self._reduce(self._value, self.unit / other)
        if isinstance(other, Quantity):
            return self._reduce(_arr(self._value) / _arr(other._value), self.unit / other.unit)
        return Quantity(_arr(self._value) / other, self.unit)

    def __rtruediv__(self, other):
        return Quantity(other / _arr(self._value), dimensionless / self.unit)

    __div__ = __truediv__
    __rdiv__ = __rtruediv__

    def __add__(self, other):
        if not isinstance(other, Quantity):
            if self.unit.is_dimensionless():
                return self._value * self.unit.factor + other
            raise TypeError('Cannot add a Quantity and a plain number')
        return Quantity(_arr(self._value) + _arr(other.value_in_unit(self.unit)), self.unit)

    __radd__ = __add__

    def __sub__(self, other):
        if not isinstance(other, Quantity):
            if self.unit.is_dimensionless():
                return self._value * self.unit.factor - other
            raise TypeError('Cannot subtract a plain number from a Quantity')
        return Quantity(_arr(self._value) - _arr(other.value_in_unit(self.unit)), self.unit)

    def __rsub__(self, other):
        return (-self).__add__(other)

    def __neg__(self):
        return Quantity(-_arr(self._value), self.unit)

    def __pos__(self):
        return self

    def __abs__(self):
        return Quantity(abs(_arr(self._value)), self.unit)

    def __pow__(self, p):
        return Quantity(_arr(self._value) ** p, self.unit ** p)

    def sqrt(self):
        return Quantity(_np.sqrt(self._value), self.unit ** 0.5)

    def sum(self, *a, **k):
        return Quantity(_np.sum(self._value, *a, **k), self.unit)

    def mean(self, *a, **k):
        return Quantity(_np.mean(self._value, *a, **k), self.unit)

    def max(self, *a, **k):
        return Quantity(_np.max(self._value, *a, **k), self.unit)

    def min(self, *a, **k):
        return Quantity(_np.min(self._value, *a, **k), self.unit)

    # -- comparisons -----------------------------------------------------------------------
    def _cmp_value(self, other):
        if isinstance(other, Quantity):
            return other.value_in_unit(self.unit)
        if self.unit.is_dimensionless():
            return other / self.unit.factor
        raise TypeError('Cannot compare a Quantity with a plain number')

    def __eq__(self, other):
        if not isinstance(other, Quantity):
            return False
        if not self.unit.is_compatible(other.unit):
            return False
        r = _arr(self._value) == _arr(other.value_in_unit(self.unit))
        return r

    def __ne__(self, other):
        r = self.__eq__(other)
        return ~r if isinstance(r, _np.ndarray) else not r

    __hash__ = None

    def __lt__(self, other):
        return _arr(self._value) < self._cmp_value(other)

    def __le__(self, other):
        return _arr(self._value) <= self._cmp_value(other)

    def __gt__(self, other):
        return _arr(self._value) > self._cmp_value(other)

    def __ge__(self, other):
        return _arr(self._value) >= self._cmp_value(other)

    def __bool__(self):
        return bool(_np.any(self._value)) if isinstance(self._value, _np.ndarray) else bool(self._value)

    # -- container behaviour ---------------------------------------------------------------
    def __len__(self):
        return len(self._value)

    def __getitem__(self, key):
        return Quantity(_arr(self._value)[key] if isinstance(key, (list, _np.ndarray)) else self._value[key],
                        self.unit)

    def __setitem__(self, key, value):
        if isinstance(value, Quantity):
            value = value.value_in_unit(self.unit)
        elif not self.unit.is_dimensionless():
            raise TypeError('Cannot assign a plain number into a Quantity container')
        self._value[key] = value

    def __iter__(self):
        for v in self._value:
            yield Quantity(v, self.unit)

    def append(self, item):
        self._value.append(item.value_in_unit(self.unit))

    def __float__(self):
        if not self.unit.is_dimensionless():
            raise TypeError('only dimensionless Quantities convert to float')
        return float(self._value) * self.unit.factor

    def __array__(self, dtype=None, copy=None):
        return _np.asarray(self._value, dtype=dtype)

    @property
    def shape(self):
        return _np.shape(self._value)

    def __str__(self):
        return '%s %s' % (self._value, self.unit.get_name())

    def __repr__(self):
        return 'Quantity(value=%r, unit=%s)' % (self._value, self.unit.get_name())

    def format(self, fmt):
        return '%s %s' % (fmt % self._value, self.unit.get_name())

    def __copy__(self):
        import copy
        return Quantity(copy.copy(self._value), self.unit)

    def __deepcopy__(self, memo):
        import copy
        return Quantity(copy.deepcopy(self._value, memo), self.unit)


def _arr(v):
    if isinstance(v, (list, tuple)):
        return _np.asarray(v, dtype=float)
    return v


def is_quantity(x):
    return isinstance(x, Quantity)


def is_unit(x):
    return isinstance(x, Unit)


# ---- base and derived units (factor = size in SI: kg, m, s, K, C, mol, rad) -------------------
_E_CHARGE = 1.602176487e-19
meter = meters = _base('meter', 1.0, length=1)
nanometer = nanometers = _base('nanometer', 1e-9, length=1)
angstrom = angstroms = _base('angstrom', 1e-10, length=1)
picometer = picometers = _base('picometer', 1e-12, length=1)
centimeter = centimeters = _base('centimeter', 1e-2, length=1)
second = seconds = _base('second', 1.0, time=1)
millisecond = milliseconds = _base('millisecond', 1e-3, time=1)
microsecond = microseconds = _base('microsecond', 1e-6, time=1)
nanosecond = nanoseconds = _base('nanosecond', 1e-9, time=1)
picosecond = picoseconds = _base('picosecond', 1e-12, time=1)
femtosecond = femtoseconds = _base('femtosecond', 1e-15, time=1)
minute = minutes = _base('minute', 60.0, time=1)
hour = hours = _base('hour', 3600.0, time=1)
day = days = _base('day', 86400.0, time=1)
kelvin = kelvins = _base('kelvin', 1.0, temperature=1)
mole = moles = _base('mole', 1.0, amount=1)
kilogram = kilograms = _base('kilogram', 1.0, mass=1)
gram = grams = _base('gram', 1e-3, mass=1)
dalton = daltons = amu = amus = _base('dalton', 1e-3, mass=1, amount=-1)
elementary_charge = elementary_charges = _base('elementary charge', _E_CHARGE, charge=1)
coulomb = coulombs = _base('coulomb', 1.0, charge=1)
radian = radians = _base('radian', 1.0, angle=1)
degree = degrees = _base('degree', math.pi / 180.0, angle=1)
joule = joules = _base('joule', 1.0, mass=1, length=2, time=-2)
kilojoule = kilojoules = _base('kilojoule', 1e3, mass=1, length=2, time=-2)
calorie = calories = _base('calorie', 4.184, mass=1, length=2, time=-2)
kilocalorie = kilocalories = _base('kilocalorie', 4184.0, mass=1, length=2, time=-2)
kilojoule_per_mole = kilojoules_per_mole = Unit(1e3, joule.dims, 'kilojoule/mole') / Unit(1.0, mole.dims, None)
kilojoule_per_mole._name = 'kilojoule/mole'
kilocalorie_per_mole = kilocalories_per_mole = Unit(4184.0, joule.dims) / Unit(1.0, mole.dims)
kilocalorie_per_mole._name = 'kilocalorie/mole'
pascal = pascals = _base('pascal', 1.0, mass=1, length=-1, time=-2)
bar = bars = _base('bar', 1e5, mass=1, length=-1, time=-2)
atmosphere = atmospheres = _base('atmosphere', 101325.0, mass=1, length=-1, time=-2)
molar = _base('molar', 1e3, amount=1, length=-3)

_MD_UNITS = [nanometer, picosecond, kelvin, elementary_charge, radian, mole, dalton]

BOLTZMANN_CONSTANT_kB = Quantity(1.3806504e-23, joule / kelvin)
AVOGADRO_CONSTANT_NA = Quantity(6.02214179e23, dimensionless / mole)
MOLAR_GAS_CONSTANT_R = BOLTZMANN_CONSTANT_kB * AVOGADRO_CONSTANT_NA


def _md_unit_for(u):
    """Unit with the same dimension as ``u`` built from nm, ps, dalton, K, e, mol, rad."""
    mass, length, time, temp, charge, amount, angle = u.dims
    # dalton carries amount^-1: mass dims are expressed through dalton*mole
    out = (dalton ** mass) * (mole ** (amount + mass)) * (nanometer ** length) * (picosecond ** time) * \
          (kelvin ** temp) * (elementary_charge ** charge) * (radian ** angle)
    return out


class _MdUnitSystem(object):
    pass


md_unit_system = _MdUnitSystem()


def sqrt(x):
    if isinstance(x, Quantity):
        return x.sqrt()
    return math.sqrt(x)


def norm(x):
    if isinstance(x, Quantity):
        return Quantity(float(_np.linalg.norm(_np.asarray(x._value, dtype=float))), x.unit)
    return float(_np.linalg.norm(x))


def dot(a, b):
    ua = a.unit if isinstance(a, Quantity) else dimensionless
    ub = b.unit if isinstance(b, Quantity) else dimensionless
    va = a._value if isinstance(a, Quantity) else a
    vb = b._value if isinstance(b, Quantity) else b
    return Quantity(_np.dot(va, vb), ua * ub)
