"""Extract the SFMT-19937 known-answer vector the reference's own test holds
(src/tests/test_random.cpp:433-509, `Random(4321)` -> nextULong()) into a fixture.
Run in the build container only (reads /root/reference); the fixture is committed."""
import re
import sys

src = open("/root/reference/src/tests/test_random.cpp").read()
body = src[src.index("static const uint64_t reference[]"):]
body = body[:body.index("};")]
words = re.findall(r"0x([0-9a-fA-F]{16})ULL", body)
with open(sys.argv[1] if len(sys.argv) > 1 else "tests/golden/sfmt_kat_seed4321.txt", "w") as f:
    f.write("# Random(4321).nextULong() x %d -- src/tests/test_random.cpp:436-503\n" % len(words))
    for w in words:
        f.write(w.lower() + "\n")
print(len(words), "words")
