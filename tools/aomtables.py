"""Read normative AV1 constant tables out of the unstripped libaom shared object that ships
with opencv-python-headless.  The tables are facts of the AV1
specification; this module only fetches their bytes so that the generated header
(av1-go_b200/csrc/tables_gen.inc) is exact.  Used by tools/gen_tables.py only.
"""
import ctypes as C
import struct

import numpy as np

from . import aomenc


class Elf:
    def __init__(self, path):
        self.path = path
        with open(path, "rb") as f:
            self.data = f.read()
        d = self.data
        assert d[:4] == b"\x7fELF" and d[4] == 2
        (self.e_phoff, self.e_shoff) = struct.unpack_from("<QQ", d, 0x20)
        (self.e_phentsize, self.e_phnum, self.e_shentsize, self.e_shnum, self.e_shstrndx) = struct.unpack_from("<HHHHH", d, 0x36)
        self.ph = []
        for i in range(self.e_phnum):
            o = self.e_phoff + i * self.e_phentsize
            p_type, p_flags, p_offset, p_vaddr, p_paddr, p_filesz, p_memsz, p_align = struct.unpack_from("<IIQQQQQQ", d, o)
            if p_type == 1:
                self.ph.append((p_vaddr, p_offset, p_filesz))
        sh = []
        for i in range(self.e_shnum):
            o = self.e_shoff + i * self.e_shentsize
            sh.append(struct.unpack_from("<IIQQQQIIQQ", d, o))
        self.syms = {}
        for s in sh:
            if s[1] == 2:  # SHT_SYMTAB
                strtab = sh[s[6]]
                stroff = strtab[4]
                n = s[5] // 24
                for j in range(n):
                    o = s[4] + j * 24
                    st_name, st_info, st_other, st_shndx, st_value, st_size = struct.unpack_from("<IBBHQQ", d, o)
                    if st_size == 0 or st_shndx == 0:
                        continue
                    e = d.index(b"\0", stroff + st_name)
                    name = d[stroff + st_name:e].decode()
                    self.syms.setdefault(name, []).append((st_value, st_size, st_info & 0xF))

    def v2o(self, vaddr):
        for va, off, fsz in self.ph:
            if va <= vaddr < va + fsz:
                return off + (vaddr - va)
        raise KeyError(hex(vaddr))

    def sym_bytes(self, name, which=0, size=None):
        lst = self.syms[name]
        va, sz, _ = lst[which]
        if size is not None:
            sz = size
        o = self.v2o(va)
        return self.data[o:o + sz]

    def arr(self, name, dtype, which=0):
        return np.frombuffer(self.sym_bytes(name, which), dtype=dtype).copy()

    def addr(self, name, which=0):
        return self.syms[name][which][0]


_ELF = None


def elf():
    global _ELF
    if _ELF is None:
        aomenc.lib()
        import importlib.util, glob, os
        spec = importlib.util.find_spec("cv2")
        base = os.path.join(os.path.dirname(os.path.dirname(spec.origin)), "opencv_python_headless.libs")
        _ELF = Elf(sorted(glob.glob(os.path.join(base, "libaom-*.so*")))[0])
    return _ELF


def load_base():
    """Runtime load address of libaom in this process (for calling non-exported functions)."""
    l = aomenc.lib()
    e = elf()
    real = C.cast(l.aom_codec_av1_cx, C.c_void_p).value
    return real - e.addr("aom_codec_av1_cx")


_RTCD_DONE = False


def call_local(name, restype, argtypes, *args):
    """Call a (possibly non-exported) libaom function by symbol-table address."""
    global _RTCD_DONE
    base = load_base()
    if not _RTCD_DONE:
        _RTCD_DONE = True
        # run-time CPU dispatch tables (av1_round_shift_array etc. are function pointers until then)
        for init in ("av1_rtcd", "aom_dsp_rtcd", "aom_scale_rtcd"):
            if init in elf().syms:
                C.CFUNCTYPE(None)(base + elf().addr(init))()
    fn = C.CFUNCTYPE(restype, *argtypes)(base + elf().addr(name))
    return fn(*args)
