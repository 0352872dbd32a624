// sim_globals.h -- TEST INFRASTRUCTURE: the dynamic shared memory of the simulated kernels (one block runs at a time)
#pragma once
alignas(128) unsigned char smem_raw[232448];
alignas(128) unsigned char ava_smem_raw[232448];
alignas(128) unsigned sh_l1c[232448 / 4];
