import ctypes as C, sys
src=open(sys.argv[1]).read().encode()
opts=[o.encode() for o in sys.argv[3:]] or [b"--gpu-architecture=sm_100a", b"--std=c++17", b"-lineinfo", b"--fmad=true"]
L=C.CDLL("libnvrtc.so.12")
prog=C.c_void_p()
assert L.nvrtcCreateProgram(C.byref(prog), src, b"k.cu", 0, None, None)==0
arr=(C.c_char_p*len(opts))(*opts)
rc=L.nvrtcCompileProgram(prog, len(opts), arr)
n=C.c_size_t(); L.nvrtcGetProgramLogSize(prog, C.byref(n)); log=C.create_string_buffer(n.value); L.nvrtcGetProgramLog(prog, log)
print("rc",rc, log.value.decode()[:2000])
L.nvrtcGetCUBINSize(prog, C.byref(n)); buf=C.create_string_buffer(n.value); L.nvrtcGetCUBIN(prog, buf)
open(sys.argv[2],'wb').write(buf.raw)
