"""Headless stand-in for the handful of pygame value types gym-highway's env code touches.

TEST INFRASTRUCTURE ONLY (see oracle/README.md): lets the *unmodified* reference at
/root/reference import and step in a container without pygame, so golden vectors can be
generated from the reference's own code.  Nothing here is on the product path.

Semantics encoded:
  * math.Vector2  - two C doubles, elementwise + - with Vector2/sequence, * scalar,
                    rotate(k*90 deg) exact (pygame special-cases multiples of 90).
  * Rect          - int fields; float assignment truncates toward zero (C int cast).
  * Rect.colliderect - strict inequalities on the truncated ints.
  * sprite.Group  - insertion ordered, iteration over a snapshot.
"""
import math as _math

QUIT = 256
KEYDOWN = 768
K_UP, K_DOWN, K_LEFT, K_RIGHT, K_SPACE, K_r, K_p = 1073741906, 1073741905, 1073741904, 1073741903, 32, 114, 112


def init():
    return (0, 0)


def quit():
    return None


class _Vector2(object):
    __slots__ = ("_x", "_y")

    def __init__(self, x=0.0, y=None):
        if y is None and hasattr(x, "__len__"):
            x, y = x[0], x[1]
        elif y is None:
            y = x
        self._x = float(x)
        self._y = float(y)

    # fields are C doubles in pygame: anything assigned is converted
    x = property(lambda s: s._x, lambda s, v: setattr(s, "_x", float(v)))
    y = property(lambda s: s._y, lambda s, v: setattr(s, "_y", float(v)))

    def __len__(self):
        return 2

    def __getitem__(self, i):
        return (self._x, self._y)[i]

    def __iter__(self):
        return iter((self._x, self._y))

    def __add__(self, o):
        return _Vector2(self._x + float(o[0]), self._y + float(o[1]))

    __radd__ = __add__

    def __iadd__(self, o):
        self._x = self._x + float(o[0])
        self._y = self._y + float(o[1])
        return self

    def __sub__(self, o):
        return _Vector2(self._x - float(o[0]), self._y - float(o[1]))

    def __isub__(self, o):
        self._x = self._x - float(o[0])
        self._y = self._y - float(o[1])
        return self

    def __mul__(self, k):
        k = float(k)
        return _Vector2(self._x * k, self._y * k)

    __rmul__ = __mul__

    def __neg__(self):
        return _Vector2(-self._x, -self._y)

    def __eq__(self, o):
        try:
            return self._x == o[0] and self._y == o[1]
        except Exception:
            return False

    def __hash__(self):
        return id(self)

    def __repr__(self):
        return "<Vector2(%r, %r)>" % (self._x, self._y)

    def rotate(self, angle):
        a = _math.fmod(float(angle), 360.0)
        if a < 0:
            a += 360.0
        eps = 1e-6
        if _math.fmod(a + eps, 90.0) < 2 * eps:
            q = int((a + eps) / 90.0) % 4
            if q == 0:
                return _Vector2(self._x, self._y)
            if q == 1:
                return _Vector2(-self._y, self._x)
            if q == 2:
                return _Vector2(-self._x, -self._y)
            return _Vector2(self._y, -self._x)
        r = _math.radians(a)
        c, s = _math.cos(r), _math.sin(r)
        return _Vector2(c * self._x - s * self._y, s * self._x + c * self._y)


class Rect(object):
    __slots__ = ("_x", "_y", "_w", "_h")

    def __init__(self, x=0, y=0, w=0, h=0):
        self._x, self._y, self._w, self._h = int(x), int(y), int(w), int(h)

    x = property(lambda s: s._x, lambda s, v: setattr(s, "_x", int(v)))
    y = property(lambda s: s._y, lambda s, v: setattr(s, "_y", int(v)))
    width = property(lambda s: s._w, lambda s, v: setattr(s, "_w", int(v)))
    height = property(lambda s: s._h, lambda s, v: setattr(s, "_h", int(v)))
    w = width
    h = height

    def colliderect(self, o):
        return (self._x < o._x + o._w and self._y < o._y + o._h
                and self._x + self._w > o._x and self._y + self._h > o._y)

    def __repr__(self):
        return "<rect(%d, %d, %d, %d)>" % (self._x, self._y, self._w, self._h)


class Surface(object):
    def __init__(self, size, *a, **k):
        self._size = (int(size[0]), int(size[1]))

    def fill(self, *a, **k):
        return None

    def set_colorkey(self, *a, **k):
        return None

    def get_rect(self, **k):
        return Rect(0, 0, self._size[0], self._size[1])

    def convert(self):
        return self

    def blit(self, *a, **k):
        return None


class _Draw(object):
    @staticmethod
    def rect(*a, **k):
        return None


class _Clock(object):
    def tick(self, *_):
        return 0

    def get_time(self):
        return 0


class _Time(object):
    Clock = _Clock


class _Event(object):
    @staticmethod
    def get():
        return []


class _Math(object):
    Vector2 = _Vector2


class _Sprite(object):
    def __init__(self, *groups):
        self._groups = {}

    def update(self, *args):
        return None


class _Group(object):
    def __init__(self, *sprites):
        self.spritedict = {}
        self.add(*sprites)

    def sprites(self):
        return list(self.spritedict)

    def add(self, *sprites):
        for s in sprites:
            if isinstance(s, _Group):
                self.add(*s.sprites())
            elif s not in self.spritedict:
                self.spritedict[s] = 0

    def remove(self, *sprites):
        for s in sprites:
            self.spritedict.pop(s, None)

    def has(self, s):
        return s in self.spritedict

    def copy(self):
        return _Group(*self.sprites())

    def update(self, *args):
        for s in self.sprites():
            s.update(*args)

    def empty(self):
        self.spritedict.clear()

    def __iter__(self):
        return iter(self.sprites())

    def __len__(self):
        return len(self.spritedict)

    def __contains__(self, s):
        return s in self.spritedict

    def __bool__(self):
        return len(self.spritedict) > 0


def _spritecollide(sprite, group, dokill, collided=None):
    hit = [s for s in group if sprite.rect.colliderect(s.rect)]
    if dokill:
        group.remove(*hit)
    return hit


class _SpriteMod(object):
    Sprite = _Sprite
    Group = _Group
    spritecollide = staticmethod(_spritecollide)


import sys as _sys
import types as _types

math = _types.ModuleType("pygame.math")
math.Vector2 = _Vector2
_sys.modules["pygame.math"] = math
Vector2 = _Vector2

sprite = _types.ModuleType("pygame.sprite")
sprite.Sprite = _Sprite
sprite.Group = _Group
sprite.spritecollide = _spritecollide
_sys.modules["pygame.sprite"] = sprite

draw = _Draw()
time = _Time()
event = _Event()
