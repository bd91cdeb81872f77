#!/usr/bin/env python3
"""TEST INFRASTRUCTURE — executes the reference's own lib/fsearch.py under CPython 3.

The reference core is RPython (translated to `fsearch-c` by bin/find_hit.py:198-208); there is
no RPython/PyPy toolchain in this image, so the translated binary cannot be built.  This shim
loads the reference *source text* from /root/reference/lib/fsearch.py at run time (nothing is
copied into this repository), registers stub `rpython.*` modules that provide the handful of
runtime services the file imports, applies eight py2->py3 / RPython-semantics textual
substitutions (none changes an algorithm) and exec()s it.

It only works inside the build container (where /root/reference is mounted).  It is used to
generate the golden vectors committed under tests/golden/ (see tests/golden/make_golden.py);
nothing in the product, the GPU tests, smoke() or bench.py imports it.
"""
import math
import mmap as _mmap
import os
import re
import struct
import sys
import types

REFERENCE_CORE = os.environ.get('SWIFTORTHO_REFERENCE', '/root/reference') + '/lib/fsearch.py'


class _MT19937:
    """rpython.rlib.rrandom.Random: MT19937 with init_genrand / genrand32 / random."""

    def __init__(self, seed=0):
        self.state = [0] * 624
        self.index = 624
        self.init_genrand(seed)

    def init_genrand(self, s):
        mt = self.state
        mt[0] = s & 0xffffffff
        for i in range(1, 624):
            mt[i] = (1812433253 * (mt[i - 1] ^ (mt[i - 1] >> 30)) + i) & 0xffffffff
        self.index = 624

    def genrand32(self):
        mt = self.state
        if self.index >= 624:
            for kk in range(624):
                y = (mt[kk] & 0x80000000) | (mt[(kk + 1) % 624] & 0x7fffffff)
                mt[kk] = mt[(kk + 397) % 624] ^ (y >> 1) ^ (0x9908b0df if y & 1 else 0)
            self.index = 0
        y = mt[self.index]
        self.index += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9d2c5680
        y ^= (y << 15) & 0xefc60000
        y ^= y >> 18
        return y & 0xffffffff

    def random(self):
        a = self.genrand32() >> 5
        b = self.genrand32() >> 6
        return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0)


class _r_uint32(int):
    def __new__(cls, v=0):
        return int.__new__(cls, int(v) & 0xffffffff)


def _intmask(v):
    v = int(v) & 0xffffffffffffffff
    return v - (1 << 64) if v >= (1 << 63) else v


class _RMmap:
    def __init__(self, fileno, length, access=None):
        size = os.fstat(fileno).st_size
        self.size = size
        self._m = _mmap.mmap(fileno, 0, access=_mmap.ACCESS_READ) if size > 0 else b''

    def getitem(self, i):
        return chr(self._m[i])

    def getslice(self, start, length):
        return self._m[start:start + length].decode('latin-1')

    def close(self):
        if self._m:
            self._m.close()


class _FileWrap:
    """open() replacement: binary file, latin-1 str in/out (py2 str semantics)."""

    def __init__(self, name, mode='r', buffering=-1):
        m = mode.replace('b', '')
        self._f = open(name, m + 'b')

    def write(self, s):
        self._f.write(s.encode('latin-1') if isinstance(s, str) else s)

    def read(self, n=-1):
        return self._f.read(n).decode('latin-1')

    def fileno(self):
        return self._f.fileno()

    def seek(self, *a):
        return self._f.seek(*a)

    def flush(self):
        self._f.flush()

    def close(self):
        self._f.close()


def _py2_str(x):
    if isinstance(x, float):
        return '%.6f' % x          # RPython str(float) == formatd(x, 'f', 6)
    return str(x)


def _install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod('rpython')
    mod('rpython.rtyper')
    mod('rpython.rtyper.lltypesystem')
    mod('rpython.rtyper.lltypesystem.module')
    mod('rpython.rtyper.lltypesystem.module.ll_math',
        ll_math_log=math.log, ll_math_log10=math.log10, ll_math_pow=math.pow)
    rrandom = mod('rpython.rlib.rrandom', Random=_MT19937)
    rfloat = mod('rpython.rlib.rfloat', erfc=math.erfc)
    mod('rpython.rtyper.lltypesystem.rffi', r_ushort=int, r_int=int)
    mod('rpython.rlib.rarithmetic', intmask=_intmask, r_uint32=_r_uint32, r_uint=int,
        string_to_int=int)
    rfile = mod('rpython.rlib.rfile')
    rmmap = mod('rpython.rlib.rmmap', mmap=_RMmap, ACCESS_READ=1, ACCESS_WRITE=2)

    class _TimSort:
        def __init__(self, lst):
            self.lst = lst

        def sort(self):
            self.lst.sort()
    listsort = mod('rpython.rlib.listsort', TimSort=_TimSort)
    rstring = mod('rpython.rlib.rstring')

    def runpack(fmt, s):
        r = struct.unpack('<' + fmt, s.encode('latin-1') if isinstance(s, str) else s)
        return r[0] if len(r) == 1 else r
    mod('rpython.rlib.rstruct')
    mod('rpython.rlib.rstruct.runpack', runpack=runpack)
    rgc = mod('rpython.rlib.rgc', collect=lambda *a: None)
    mod('rpython.rlib', rrandom=rrandom, rfloat=rfloat, rfile=rfile, rmmap=rmmap,
        listsort=listsort, rstring=rstring, rgc=rgc)


def load_reference():
    """Return the exec'd globals of the reference core (functions: entry_point, kswat_st, seg, ...)."""
    _install_stubs()
    src = open(REFERENCE_CORE, encoding='latin-1').read()
    # 1. print statements -> functions
    src = re.sub(r'^(\s*)print (.+)$', r'\1print(\2)', src, flags=re.M)
    # 2. xrange
    src = src.replace('xrange', 'range')
    # 3. py2 integer division in guess_start (fsearch.py:2549 and the dead twin :1921)
    src = src.replace('dist /= N', 'dist //= N')
    # 4. float slice bound is cast to int by RPython (fsearch.py:3062)
    src = src.replace('hits[:vmax]', 'hits[:int(vmax)]')
    # 5. '%f' % int prints the int in RPython (fsearch.py:3242): the 12th field is the bit score
    old = "'%s\\t%s\\t%s\\t%d\\t%d\\t%d\\t%d\\t%d\\t%d\\t%d\\t%s\\t%f\\t%d\\t%d\\t%d\\t%s\\n' % ("
    new = "'%s\\t%s\\t%s\\t%d\\t%d\\t%d\\t%d\\t%d\\t%d\\t%d\\t%s\\t%s\\t%d\\t%d\\t%d\\t%s\\n' % ("
    assert src.count(old) == 1
    src = src.replace(old, new)
    # 6. range object is immutable in py3 (fsearch.py:408)
    src = src.replace('aa_nr_tbl = range(512)', 'aa_nr_tbl = list(range(512))')
    # 7. invalid line in dead code `kolmogorov` (fsearch.py:2820)
    src = re.sub(r'^(\s*)n = len\(S\), .*$', r'\1n = len(S)', src, flags=re.M)
    # 8. AL == 0: RPython float division yields NaN, CPython raises (fsearch.py:1471)
    src = src.replace('idy *= (100. / AL)', "idy = idy * (100. / AL) if AL else float('nan')")
    g = {'__name__': 'fsearch_reference', 'open': _FileWrap, 'str': _py2_str}
    exec(compile(src, REFERENCE_CORE, 'exec'), g)
    return g


def run_entry_point(argv):
    g = load_reference()
    return g['entry_point'](['fsearch-c'] + list(argv))


if __name__ == '__main__':
    sys.setrecursionlimit(100000)
    run_entry_point(sys.argv[1:])
