"""Developer tool: per-kernel SASS mnemonic counts of the shipped library (second half of profiles/rNN_sass_evidence.txt).
    python tools/sass_per_kernel.py [lib.so]"""
import re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "3dfacerecon_b200/lib3dfacerecon_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
blocks = sass.split("Function : ")[1:]
cols = [("UTCHMMA", r"\bUTCHMMA\b"), ("LDTM", r"\bLDTM\b"), ("UBLKCP", r"\bUBLKCP\b"), ("REDG", r"\bREDG\b"), ("VIMNMX3", r"\bVIMNMX3\b"),
        ("SYNCS", r"\bSYNCS\b"), ("ACQBULK", r"\bACQBULK\b"), ("PREEXIT", r"\bPREEXIT\b")]
print("".join("%8s" % c for c, _ in cols) + "  instrs  kernel")
for name, blk in zip(names, blocks):
    body = blk.split("\n", 1)[1]
    n = len(re.findall(r"^\s+/\*[0-9a-f]{4}\*/", body, re.M))
    short = re.sub(r"\(.*", "", name)
    print("".join("%8d" % len(re.findall(rx, body)) for _, rx in cols) + "  %6d  %s" % (n, short))
