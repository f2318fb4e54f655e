#!/usr/bin/env python
"""md5 of every kernel's SASS (addresses and encodings stripped) in the built objects of lqrrt_b200/csrc/build:
a CPU-only regression guard -- a change that is meant to leave a kernel alone must leave its fingerprint alone.

    python tools/sass_fingerprint.py > /tmp/before.txt ; (edit, rebuild) ; python tools/sass_fingerprint.py | diff /tmp/before.txt -
"""
import glob
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    pat = sys.argv[1] if len(sys.argv) > 1 else ''
    for obj in sorted(glob.glob(os.path.join(ROOT, 'lqrrt_b200', 'csrc', 'build', '*.o'))):
        out = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
        for blk in out.split('Function : ')[1:]:
            name = blk.split('\n', 1)[0].strip()
            if pat and pat not in name:
                continue
            demangled = subprocess.run(['c++filt', name], capture_output=True, text=True).stdout.strip()
            body = []
            for ln in blk.split('\n'):
                m = re.match(r'\s+/\*[0-9a-f]{4}\*/\s+(.*?)\s*/\*', ln)
                if m:
                    body.append(re.sub(r'0x[0-9a-f]+', 'ADDR', m.group(1)) if 'BRA' in m.group(1) or 'CALL' in m.group(1) else m.group(1))
            print("%s %5d %s :: %s" % (hashlib.md5('\n'.join(body).encode()).hexdigest()[:12], len(body), os.path.basename(obj),
                                       re.sub(r'\(.*', '', demangled)[:150]))


if __name__ == '__main__':
    main()
