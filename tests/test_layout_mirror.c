/* C mirror of the struct layouts the Rust shim (rust/gpu.rs) and the C++ mirror cast through.
 * Compiled (not run) by tests/test_abi.py; it pins the byte sizes include/b200msm.h documents. */
#include <stddef.h>
#include <stdint.h>
#include "b200msm.h"
typedef struct { uint64_t l[6]; } blst_fp;              /* reference src/fp.rs:482-491 */
typedef struct { blst_fp fp[2]; } blst_fp2;             /* reference src/fp2.rs:450-454 */
typedef struct { blst_fp x, y; } blst_p1_affine;        /* G1Affine   src/g1.rs:54-56   */
typedef struct { blst_fp x, y, z; } blst_p1;            /* G1Projective src/g1.rs:435-437 */
typedef struct { blst_fp2 x, y; } blst_p2_affine;       /* G2Affine   src/g2.rs:66-68   */
typedef struct { blst_fp2 x, y, z; } blst_p2;           /* G2Projective src/g2.rs:415-417 */
typedef struct { uint64_t l[4]; } blst_fr;              /* Scalar     src/scalar.rs:23-25 */
_Static_assert(sizeof(blst_p1_affine) == 96 && _Alignof(blst_p1_affine) == 8, "G1Affine");
_Static_assert(sizeof(blst_p1) == 144, "G1Projective = out[18]");
_Static_assert(sizeof(blst_p2_affine) == 192, "G2Affine");
_Static_assert(sizeof(blst_p2) == 288, "G2Projective = out[36]");
_Static_assert(sizeof(blst_fr) == 32, "Scalar / BigInt<4>");
_Static_assert(offsetof(blst_p1_affine, y) == 48 && offsetof(blst_p2_affine, y) == 96, "x then y");
int main(void) { return 0; }
