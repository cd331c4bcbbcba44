"""include/ovp.h must be a plain C header (the FFI boundary a reference maintainer binds): compile a C and a C++ translation unit
against it and link them with the library's exported symbols (no GPU needed)."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

C_SRC = r'''
#include "ovp.h"
#include <stdio.h>
int main(void) {
  ovp_state_options o = {1, 0, 1, 1, 1, 1, 11, 1024, 0.01, 1.0, 1.0, 0.01, 0.75, 1.0};
  ovp_feature_batch b; ovp_updater_options u = {1.0, 1.0};
  (void)o; (void)b; (void)u;
  printf("%s %d\n", ovp_status_string(OVP_ERR_NON_CONTIGUOUS), (int)sizeof(ovp_feature_batch));
  return 0;
}
'''


def test_header_compiles_as_c_and_cpp_and_links():
    from ov_plane_b200 import api
    if not os.path.exists(api.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    libdir = os.path.dirname(api.LIB_PATH)
    with tempfile.TemporaryDirectory() as d:
        for comp, ext in (("gcc", "c"), ("g++", "cpp")):
            src = os.path.join(d, "t." + ext)
            open(src, "w").write(C_SRC)
            exe = os.path.join(d, "t_" + ext)
            subprocess.check_call([comp, "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-o", exe, "-L", libdir, "-lovp",
                                   "-Wl,-rpath," + libdir, "-Wl,--allow-shlib-undefined"])
            out = subprocess.check_output([exe]).decode()
            assert "non-contiguous" in out
