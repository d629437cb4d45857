/* Plain-C client of the metalens_b200 C-ABI (include/metalens_b200.h): the calls below are host-only, so the program
 * runs without a GPU.  tests/test_abi.py compiles it with gcc, links it against libmetalens_b200.so and runs it --
 * the header is C99, the entry points have C linkage, errors come back as codes + mlb_last_error().
 *   gcc -std=c99 -I include examples/c_abi_host_only.c -L metalens_b200 -lmetalens_b200 -Wl,-rpath,$PWD/metalens_b200 */
#include <stdio.h>
#include <string.h>

#include "metalens_b200.h"

int main(void) {
    int sizes[2] = {0, 0}, radix[8], pad = 0, stages, i;
    if (mlb_version() < 100) { fprintf(stderr, "unexpected version %d\n", mlb_version()); return 1; }
    if (mlb_struct_sizes(sizes) != MLB_OK) { fprintf(stderr, "%s\n", mlb_last_error()); return 1; }
    if (sizes[0] != (int)sizeof(mlb_table_pack) || sizes[1] != (int)sizeof(mlb_lens_desc)) {
        fprintf(stderr, "descriptor layout differs: library %d/%d, header %d/%d\n", sizes[0], sizes[1],
                (int)sizeof(mlb_table_pack), (int)sizeof(mlb_lens_desc));
        return 1;
    }
    /* 3375 = 3^3 5^3: the reference's good_fft_number() grid of the bench (nearfield.py:30-36) */
    stages = mlb_fft_mixed_plan(3375, radix, &pad);
    if (stages != 3 || radix[0] != 15 || radix[1] != 15 || radix[2] != 15 || pad != 30) return 1;
    if (mlb_fft_mixed_compiled(3375, 0) != 1 || mlb_fft_mixed_compiled(225, 1) != 1) return 1;
    /* errors: a code and a thread-local message, never an abort */
    if (mlb_fft_mixed_plan(14, radix, &pad) >= 0 || strstr(mlb_last_error(), "2^a 3^b 5^c") == NULL) return 1;
    if (mlb_set_option("no_such_option", 1) == MLB_OK || mlb_get_option("no_such_option") != -1) return 1;
    printf("metalens_b200 C-ABI %d: table pack %d B, lens descriptor %d B, max FFT length %d, 3375 =", mlb_version(),
           sizes[0], sizes[1], mlb_fft_max_length());
    for (i = 0; i < stages; ++i) printf(" %d", radix[i]);
    printf("\n");
    return 0;
}
