"""SASS op-count table of the built library (cuobjdump -sass): python tools/sass_opcounts.py > profiles/<name>.csv"""
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "rga3-release_b200", "libb200vit.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "HMMA", "LDSM", "MUFU.EX2"]
print("kernel," + ",".join(OPS))
for (mangled, body), name in zip(re.findall(r"Function : (\S+)(.*?)(?=Function : |\Z)", sass, re.S), names):
    name = re.sub(r"\((int|bool|unsigned int)\)", "", name)
    name = re.sub(r"\(.*", "", name).replace("b200::<unnamed>::", "").replace("void ", "").replace(",", ";")
    c = [len(re.findall(r"\b" + re.escape(op) + r"\b" if "." not in op else re.escape(op), body)) for op in OPS]
    c[0] = len(re.findall(r"\bUTCHMMA\b", body))
    print(name + "," + ",".join(map(str, c)))
