"""Disassemble functions of a CPython-3.6 .pyc (the reference ships __pycache__/losses.cpython-36.pyc without its source).

    python tools/pyc36_dis.py _flowgradloss _flowconsist      (build container only: reads /root/reference)

A minimal unmarshaller for the 3.6 format (FLAG_REF aware) and the 3.6 wordcode opcode names that the loss functions use; the
formulas in deep_video_interpolation_extrapolation_b200/losses.py and csrc/fwb_loss.cuh were read off this output."""
import struct, sys
data = open('/root/reference/__pycache__/losses.cpython-36.pyc','rb').read()[12:]
pos = 0; refs = []
class Code:
    pass
def r8():
    global pos; v = data[pos]; pos += 1; return v
def r32():
    global pos; v = struct.unpack('<i', data[pos:pos+4])[0]; pos += 4; return v
def rbytes(n):
    global pos; v = data[pos:pos+n]; pos += n; return v
def load():
    global pos
    b = r8(); flag = b & 0x80; t = chr(b & 0x7f)
    idx = None
    if flag and t not in 'r':
        idx = len(refs); refs.append(None)
    def done(v):
        if idx is not None: refs[idx] = v
        return v
    if t == '0': return None
    if t == 'N': return done(None)
    if t == 'T': return done(True)
    if t == 'F': return done(False)
    if t == '.': return done(Ellipsis)
    if t == 'i': return done(r32())
    if t == 'g': return done(struct.unpack('<d', rbytes(8))[0])
    if t == 'f': n = r8(); return done(float(rbytes(n)))
    if t == 'l':
        n = r32(); digs = [struct.unpack('<H', rbytes(2))[0] for _ in range(abs(n))]
        v = sum(d << (15*i) for i, d in enumerate(digs)); return done(-v if n < 0 else v)
    if t in 's': n = r32(); return done(rbytes(n))
    if t in 'ut': n = r32(); return done(rbytes(n).decode('utf8', 'replace'))
    if t in 'aA': n = r32(); return done(rbytes(n).decode('latin1'))
    if t in 'zZ': n = r8(); return done(rbytes(n).decode('latin1'))
    if t == ')': n = r8(); v = tuple(load() for _ in range(n)); return done(v)
    if t == '(': n = r32(); v = tuple(load() for _ in range(n)); return done(v)
    if t == '[': n = r32(); v = [load() for _ in range(n)]; return done(v)
    if t == '<' or t == '>': n = r32(); v = frozenset(load() for _ in range(n)); return done(v)
    if t == 'r': return refs[r32()]
    if t == 'c':
        c = Code()
        c.argcount, c.kwonly, c.nlocals, c.stacksize, c.flags = r32(), r32(), r32(), r32(), r32()
        c.code = load(); c.consts = load(); c.names = load(); c.varnames = load(); c.freevars = load(); c.cellvars = load()
        c.filename = load(); c.name = load(); c.firstlineno = r32(); c.lnotab = load()
        return done(c)
    raise ValueError('type %r at %d' % (t, pos))
top = load()
# python 3.6 opcode names (subset), wordcode
op = {1:'POP_TOP',2:'ROT_TWO',3:'ROT_THREE',4:'DUP_TOP',10:'UNARY_POSITIVE',11:'UNARY_NEGATIVE',12:'UNARY_NOT',15:'UNARY_INVERT',19:'BINARY_POWER',20:'BINARY_MULTIPLY',22:'BINARY_MODULO',23:'BINARY_ADD',24:'BINARY_SUBTRACT',25:'BINARY_SUBSCR',26:'BINARY_FLOOR_DIVIDE',27:'BINARY_TRUE_DIVIDE',55:'INPLACE_ADD',56:'INPLACE_SUBTRACT',57:'INPLACE_MULTIPLY',68:'GET_ITER',83:'RETURN_VALUE',87:'POP_BLOCK',90:'STORE_NAME',92:'UNPACK_SEQUENCE',93:'FOR_ITER',95:'STORE_ATTR',100:'LOAD_CONST',101:'LOAD_NAME',102:'BUILD_TUPLE',103:'BUILD_LIST',106:'LOAD_ATTR',107:'COMPARE_OP',108:'IMPORT_NAME',110:'JUMP_FORWARD',113:'JUMP_ABSOLUTE',114:'POP_JUMP_IF_FALSE',115:'POP_JUMP_IF_TRUE',116:'LOAD_GLOBAL',120:'SETUP_LOOP',124:'LOAD_FAST',125:'STORE_FAST',131:'CALL_FUNCTION',132:'MAKE_FUNCTION',133:'BUILD_SLICE',135:'LOAD_CLOSURE',136:'LOAD_DEREF',137:'STORE_DEREF',141:'CALL_FUNCTION_KW',142:'CALL_FUNCTION_EX',160:'LOAD_METHOD',161:'CALL_METHOD'}
def dis(c, want):
    for k in c.consts:
        if isinstance(k, Code): dis(k, want)
    if not any(w in c.name for w in want): return
    print('====', c.name, 'args', c.varnames[:c.argcount], 'line', c.firstlineno)
    code = c.code
    for i in range(0, len(code), 2):
        o, a = code[i], code[i+1]
        nm = op.get(o, str(o)); ex = ''
        if nm in ('LOAD_CONST',): ex = repr(c.consts[a]) if not isinstance(c.consts[a], Code) else '<code %s>' % c.consts[a].name
        elif nm in ('LOAD_GLOBAL','LOAD_ATTR','LOAD_NAME','STORE_ATTR','LOAD_METHOD','IMPORT_NAME','STORE_NAME'): ex = c.names[a]
        elif nm in ('LOAD_FAST','STORE_FAST'): ex = c.varnames[a]
        print('  %3d %-20s %3d %s' % (i, nm, a, ex))
dis(top, sys.argv[1:])
