"""CPU: the C-ABI library builds for sm_100a, loads, and exports exactly what include/d3f.h declares;
host-side argument checking works without a GPU (no compute calls here)."""
import ctypes
import os
import re
import subprocess

import pytest

from d3fields_b200 import _native, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'd3f.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(d3f_[a-z_0-9]+)\s*\(', src)))


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(_native.SYMBOLS)


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    for s in _declared_symbols():
        assert hasattr(lib, s), f'{s} declared in include/d3f.h but not exported by libd3f.so'
    out = subprocess.run(['nm', '-D', '--defined-only', path], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r' T (d3f_[a-z_0-9]+)', out)))
    assert exported == _declared_symbols()


def test_library_is_sm100a_only_and_has_no_torch_dependency():
    path = build.build()
    out = subprocess.run(['cuobjdump', '--list-elf', path], capture_output=True, text=True).stdout
    archs = set(re.findall(r'sm_(\d+a?)', out))
    assert archs == {'100a'}, archs
    ldd = subprocess.run(['ldd', path], capture_output=True, text=True).stdout
    assert 'torch' not in ldd and 'c10' not in ldd


def test_abi_version_and_argument_validation_without_gpu():
    lib = _native.load()
    assert lib.d3f_abi_version() == _native.ABI_VERSION
    obs = _native.D3FObs(0, 480, 640, 1, 1, 1)        # V = 0 is rejected before any CUDA call
    rc = lib.d3f_eval(ctypes.byref(obs), 1, 10, None, 0, 1, 1, None, None, 0, 0.02, None)
    assert rc == -1 and b'V=0' in lib.d3f_last_error()
    obs = _native.D3FObs(4, 480, 640, 1, 1, 1)
    rc = lib.d3f_eval(ctypes.byref(obs), 1, 10, None, 0, 1, 1, None, None, 0, -1.0, None)
    assert rc == -1 and b'mu' in lib.d3f_last_error()
    rc = lib.d3f_eval(ctypes.byref(obs), 1, 10, None, 3, 1, 1, None, None, 0, 0.02, None)
    assert rc == -1 and b'keys' in lib.d3f_last_error()
    rc = lib.d3f_eval(ctypes.byref(obs), 1, 10, None, 0, 1, 1, None, None, 64, 0.02, None)
    assert rc == -1 and b'flag' in lib.d3f_last_error()
    with pytest.raises(_native.D3FError):
        _native.eval_device(99, 4, 4, 1, 1, 1, 1, 1, [], 1, 1, [], None, 0, 0.02, 0)


def test_kernel_resource_contract():
    """The walk's performance rests on compile-time properties that can be checked without a GPU: the wide tile
    kernel keeps its 64-register texel cache without spilling at <= 128 registers (2 CTAs/SM), the light
    instantiation fits 4 CTAs/SM (<= 64 registers), and the hot loop uses the uniform mask reduction (REDUX), 128-bit
    streaming stores and L1 evict-last texel loads."""
    build.build()
    log = open(os.path.join(build.LIB_DIR, 'build.log')).read()
    if 'field_tile_kernel' not in log:                      # library was current: rebuild once to get ptxas -v output
        build.build(force=True)
        log = open(os.path.join(build.LIB_DIR, 'build.log')).read()
    blocks = re.findall(r"Function properties for (\S*field_tile_kernel\S*)\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, "
                        r"(\d+) bytes spill loads\nptxas info\s*: Used (\d+) registers", log)
    assert len(blocks) >= 28, 'expected 28 instantiations of field_tile_kernel in the ptxas log'
    for name, stack, st, ld, regs in blocks:
        wide = re.search(r'field_tile_kernelILb[01]ELi\d+ELb1ELb[01]ELi[0148]EEE', name) is not None
        assert int(st) == 0 and int(ld) == 0, f'{name}: spills'
        assert int(regs) <= (128 if wide else 64), f'{name}: {regs} registers'
    # the one-launch tracking iteration keeps 16 texel rows (64 registers) per thread: 2 CTAs of 256 threads per SM without
    # spilling, and a 3-CTA budget (<= 85 registers) that may spill a few words
    steps = re.findall(r"Function properties for (\S*track_step_kernelILb[01]ELi([23])\S*)\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, "
                       r"(\d+) bytes spill loads\nptxas info\s*: Used (\d+) registers", log)
    assert len(steps) == 4, steps
    for name, minb, stack, st, ld, regs in steps:
        if minb == '2':
            assert int(st) == 0 and int(ld) == 0 and int(regs) <= 128, (name, st, ld, regs)
        else:
            assert int(regs) <= 85 and int(st) <= 128 and int(ld) <= 128, (name, st, ld, regs)
    sass = subprocess.run(['cuobjdump', '-sass', build.LIB_PATH], capture_output=True, text=True).stdout
    tile = sass[sass.index('field_tile_kernelILb0ELi0ELb1ELb0ELi4'):]
    tile = tile[:tile.index('Function :', 10)] if 'Function :' in tile[10:] else tile
    assert 'REDUX.OR' in tile, 'uniform mask reduction missing'
    assert 'STG.E.EF.128' in tile, '128-bit streaming (evict-first) row stores missing'
    assert tile.count('LDG.E.EL.128.CONSTANT') >= 16, 'L1 evict-last 128-bit texel loads missing'


def test_integration_stub_matches_the_struct_layouts():
    """INTEGRATION.md shows the ctypes structs a maintainer of the reference would paste into fusion.py: their field
    names and order must be the binding's (and therefore the header's)."""
    md = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    def fields(cls):
        body = md[md.index(f'class {cls}(C.Structure)'):]
        body = body[:body.index('\n\n')]
        return re.findall(r"\('([a-zA-Z_]+)',", body)
    assert fields('_Obs') == [n for n, _ in _native.D3FObs._fields_]
    assert fields('_Key') == [n for n, _ in _native.D3FKey._fields_]
    hdr = open(os.path.join(ROOT, 'include', 'd3f.h')).read()
    key = hdr[hdr.index('typedef struct D3FKey {'):hdr.index('} D3FKey;')]
    key = re.sub(r'/\*.*?\*/', '', key, flags=re.S)
    names = re.findall(r'\b([a-zA-Z_]+)\s*(?:,|;)', key)
    assert names == [n for n, _ in _native.D3FKey._fields_], names


def test_track_step_eligibility_and_validation_without_gpu():
    """d3f_track_step_supported is host logic: the one-launch tracking iteration takes V <= 4 and a float32 map with
    C % 4 == 0, C <= 1024, strides multiples of 4 and a 16-byte aligned base; everything else must say no (the tracker then
    uses the four-launch iteration), and d3f_track_step itself must refuse it before any CUDA call."""
    F32, U8 = 0, 1
    ok = lambda V, key: _native.track_step_supported(V, 480, 640, 16, 16, 16, key)
    assert ok(4, (256, F32, 48, 64, 1024))
    assert ok(1, (256, F32, 48, 64, 4))
    assert ok(4, (256, F32, 48, 64, 256, None, (48 * 80 * 256, 80 * 256, 256)))       # a padded map
    assert not ok(5, (256, F32, 48, 64, 1024))                                       # more views than register slots
    assert not ok(4, (256, F32, 48, 64, 1028))                                       # wider than the CTA
    assert not ok(4, (256, F32, 48, 64, 6))                                          # not float4 rows
    assert not ok(4, (256, U8, 48, 64, 64))                                          # byte maps
    assert not ok(4, (260, F32, 48, 64, 64))                                         # 4-byte aligned base only
    assert not ok(4, (256, F32, 48, 64, 64, None, (48 * 64 * 64 + 2, 64 * 64, 64)))  # stride not a multiple of 4
    lib = _native.load()
    with pytest.raises(_native.D3FError):
        _native.track_step(5, 480, 640, 16, 16, 16, (256, F32, 48, 64, 1024), 16, 100.0, 16, 16, None, 0, 0.02, 0,
                           t_in=16, r_in=16, t_out=32, r_out=32, m_t=16, v_t=16, m_r=16, v_r=16, last_pts=16,
                           grad_pts=None, pts=None, n_inst=1, n_pts=10, step=1.0, lr=0.01, beta1=0.9, beta2=0.999,
                           eps=1e-8, reg_w=1.0)
    assert b'track_step' in lib.d3f_last_error()
    with pytest.raises(_native.D3FError):                                            # outputs aliasing the inputs
        _native.track_step(4, 480, 640, 16, 16, 16, (256, F32, 48, 64, 1024), 16, 100.0, 16, 16, None, 0, 0.02, 0,
                           t_in=16, r_in=16, t_out=16, r_out=32, m_t=16, v_t=16, m_r=16, v_r=16, last_pts=16,
                           grad_pts=None, pts=None, n_inst=1, n_pts=10, step=1.0, lr=0.01, beta1=0.9, beta2=0.999,
                           eps=1e-8, reg_w=1.0)
    assert b'alias' in lib.d3f_last_error()
