#!/usr/bin/env python3
"""slang2cpp.py -- TEST INFRASTRUCTURE (oracle/). Transliterates the reference's Slang shader sources into C++ so that g++ can
compile the reference's OWN closure / sampling / colour code into oracle/_ref/ (git-ignored; the reference sources are read where
they lie under /root/reference and are never copied into the repository).

    slang2cpp.py <reference>/src/shaders <out.inc> file1.slang entry.slang=entryName ...

Steps: (1) the C preprocessor resolves the #include graph of the listed files (-DVKRT_SHADER, as the reference's own build passes to
slangc: src/shaders/meson.build:69-92); (2) a token-level rewrite of the few Slang constructs that are not C++:
    [[vk::binding]] / [vk::image_format] / [shader] / [unroll] / [mutating] / [numthreads]   -> dropped
    `inout T x`, `out T x`, `in T x`                                                        -> `T& x`, `T& x`, `T x`
    `__init(...)` inside `struct S`                                                         -> `S(...)` (+ `S() = default;`)
    `this.`                                                                                 -> `this->`
    vector swizzles `.xyz`, `.xy`, `.zw`, `.rgb` read as values                             -> `.xyz()` ...
    `v.xy *= s;`  (the one swizzle store in the tree, material/textures.slang:149)          -> per-component stores
    `float2(rand(rng), rand(rng))` (arguments that advance the RNG)                          -> `float2{rand(rng), rand(rng)}`: Slang
                                            evaluates arguments left to right, C++ only guarantees that for braced initialisers
    floating literals without suffix (`1.0`, `1e-6`)                                        -> `1.0f`, `1e-6f` (Slang literals are fp32)
No expression is re-ordered and no arithmetic is changed: the output computes what the Slang source says, in fp32, with the HLSL
intrinsics supplied by hlsl_prelude.h."""
import re
import subprocess
import sys


def preprocess(shader_root, files):
    # `path.slang=name`: an entry-point file; its `main` becomes `name` (every shader stage of the reference is called main)
    unity = ""
    for f in files:
        if "=" in f:
            path, name = f.split("=")
            unity += '#define main %s\n#include "%s"\n#undef main\n' % (name, path)
        else:
            unity += '#include "%s"\n' % f
    out = subprocess.run(["cpp", "-P", "-undef", "-nostdinc", "-x", "c", "-DVKRT_SHADER", "-I", shader_root, "-"], input=unity.encode(),
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=False)
    if out.returncode != 0:
        sys.stderr.write(out.stderr.decode())
        raise SystemExit("cpp failed")
    return out.stdout.decode()


ATTRS = [r"\[\[vk::binding\([^)]*\)\]\]", r"\[vk::image_format\(\"[^\"]*\"\)\]", r"\[shader\(\"[^\"]*\"\)\]", r"\[unroll\]", r"\[mutating\]",
         r"\[numthreads\([^)]*\)\]", r"\[ForceInline\]", r"\[noinline\]"]


def match_brace(text, open_pos):
    depth = 0
    for i in range(open_pos, len(text)):
        c = text[i]
        if c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                return i
    raise ValueError("unbalanced braces")


def rewrite_structs(text):
    out = []
    pos = 0
    for m in re.finditer(r"\bstruct\s+([A-Za-z_]\w*)\s*\{", text):
        if m.start() < pos:
            continue
        name = m.group(1)
        open_pos = m.end() - 1
        close_pos = match_brace(text, open_pos)
        body = text[open_pos + 1:close_pos]
        if "__init" in body:
            has_default = re.search(r"__init\s*\(\s*\)", body) is not None
            body = re.sub(r"\b__init\s*\(", name + "(", body)
            if not has_default:
                body = "\n    " + name + "() = default;" + body
        out.append(text[pos:open_pos + 1])
        out.append(body)
        pos = close_pos
    out.append(text[pos:])
    return "".join(out)


def expr_start(text, dot):
    """Start of the postfix expression that ends right before text[dot] == '.' (identifier / call / index / member chain)."""
    i = dot - 1
    while i >= 0:
        c = text[i]
        if c in ")]":
            close, open_ = c, "(" if c == ")" else "["
            depth = 0
            while i >= 0:
                if text[i] == close:
                    depth += 1
                elif text[i] == open_:
                    depth -= 1
                    if depth == 0:
                        break
                i -= 1
            i -= 1
            continue
        if c.isalnum() or c == "_":
            while i >= 0 and (text[i].isalnum() or text[i] == "_"):
                i -= 1
            if i >= 0 and text[i] == ".":
                i -= 1
                continue
            if i >= 1 and text[i - 1:i + 1] == "->":
                i -= 2
                continue
            break
        break
    return i + 1


BROADCAST = re.compile(r"\.(x{2,4}|y{2,4}|z{2,4}|r{2,4}|g{2,4}|b{2,4})\b(?!\s*\()")


def rewrite_broadcast_swizzles(text):
    """`expr.xxx` where expr may be a scalar: -> swz_xxx(expr) (overloaded in the entry file for float and the vector types)."""
    while True:
        m = BROADCAST.search(text)
        if not m:
            return text
        s = expr_start(text, m.start())
        text = text[:s] + "swz_" + m.group(1) + "(" + text[s:m.start()] + ")" + text[m.end():]


FUNC_HEAD = re.compile(r"^([A-Za-z_][\w<>]*(?:\s*&)?)\s+([A-Za-z_]\w*)\s*\((.*)\)$", re.S)


def split_module(text):
    """Slang resolves names module-wide; C++ needs declarations first. Cut the file-scope text into items and return
    (struct names, function prototypes, non-function items [structs, constants, resource globals] in source order, function definitions)."""
    structs = re.findall(r"\bstruct\s+([A-Za-z_]\w*)\s*\{", text)
    protos, decls, funcs = [], [], []
    depth = 0
    start = 0
    head_is_func = False
    for i, c in enumerate(text):
        if c == "{":
            if depth == 0:
                head = text[start:i].strip()
                m = FUNC_HEAD.match(head)
                head_is_func = bool(m) and not head.startswith("struct") and m.group(1) not in ("return", "else")
                if head_is_func:
                    protos.append(re.sub(r"\s+", " ", head) + ";")
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0 and head_is_func:
                funcs.append(text[start:i + 1].strip())
                start = i + 1
                head_is_func = False
        elif c == ";" and depth == 0:
            item = text[start:i + 1].strip()
            if item and item != ";":
                decls.append(item)
            start = i + 1
    return structs, protos, decls, funcs


def split_args(inner):
    args, depth, cur = [], 0, ""
    for c in inner:
        if c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
        if c == "," and depth == 0:
            args.append(cur)
            cur = ""
        else:
            cur += c
    args.append(cur)
    return args


def order_argument_evaluation(text):
    """Slang evaluates call arguments left to right; C++ leaves the order unspecified (g++ goes right to left). Every argument list
    in which two or more arguments advance the RNG must therefore keep its order: vector constructors become braced initialisers
    (left-to-right by the C++ standard); anything else is refused so that a silent re-ordering cannot happen."""
    out = []
    i = 0
    n = len(text)
    while i < n:
        m = re.compile(r"\b([A-Za-z_]\w*)\s*\(").search(text, i)
        if not m:
            out.append(text[i:])
            break
        open_pos = m.end() - 1
        depth, j = 0, open_pos
        while j < n:
            if text[j] == "(":
                depth += 1
            elif text[j] == ")":
                depth -= 1
                if depth == 0:
                    break
            j += 1
        inner = text[open_pos + 1:j]
        args = split_args(inner)
        stateful = [a for a in args if re.search(r"\brng\b", a)]
        name = m.group(1)
        if len(stateful) >= 2 and name not in ("if", "for", "while", "switch", "return"):
            # a declaration `T f(inout uint rng, ...)` names rng once; a call passes it in several arguments
            if re.fullmatch(r"float[234]|uint[234]|int[234]", name):
                out.append(text[i:open_pos] + "{")
                out.append(order_argument_evaluation(inner))
                out.append("}")
                i = j + 1
                continue
            raise SystemExit("slang2cpp: argument list of %s(...) advances the RNG in %d arguments; evaluation order would be "
                             "unspecified in C++: %s" % (name, len(stateful), inner.strip()[:120]))
        out.append(text[i:open_pos + 1])
        i = open_pos + 1
    return "".join(out)


FLOAT_LIT = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+)(?![\w.])")
SWIZZLE = re.compile(r"\.((?:[xyzw]{2,4})|(?:[rgba]{2,4}))\b(?!\s*\()")


def translate(text):
    for a in ATTRS:
        text = re.sub(a, "", text)
    # the single swizzle store of the tree
    text = re.sub(r"\b(\w+)\.xy\s*\*=\s*([^;]+);", r"\1.x *= (\2); \1.y *= (\2);", text)
    text = re.sub(r"\b(?:inout|out)\s+([A-Za-z_][\w<>]*)\s+(?=[A-Za-z_])", r"\1& ", text)
    text = re.sub(r"(?<=[(,])\s*in\s+([A-Za-z_][\w<>]*)\s+(?=[A-Za-z_])", r" \1 ", text)
    text = rewrite_structs(text)
    text = re.sub(r"\bthis\.", "this->", text)
    text = FLOAT_LIT.sub(lambda m: m.group(1) + "f", text)
    text = order_argument_evaluation(text)
    text = rewrite_broadcast_swizzles(text)
    text = SWIZZLE.sub(lambda m: "." + m.group(1) + "()", text)
    return text


def main():
    if len(sys.argv) < 4:
        raise SystemExit(__doc__)
    shader_root, out_path, files = sys.argv[1], sys.argv[2], sys.argv[3:]
    text = translate(preprocess(shader_root, files))
    structs, protos, decls, funcs = split_module(text)
    with open(out_path, "w") as f:
        f.write("// GENERATED by oracle/ref_slang/slang2cpp.py from the reference's src/shaders -- do not commit, do not edit\n")
        f.write("// ---- forward declarations (Slang resolves names module-wide) ----\n")
        for n in structs:
            f.write("struct %s;\n" % n)
        for p in protos:
            f.write(p + "\n")
        f.write("// ---- types, constants and resource bindings, in source order ----\n")
        for d in decls:
            f.write(d + "\n")
        f.write("// ---- functions, in source order ----\n")
        for fn in funcs:
            f.write(fn + "\n\n")


if __name__ == "__main__":
    main()
