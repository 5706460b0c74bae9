# species file for the tabulated-opacity variant of configs[3] (our own deck): one species, opacities from a reference .aiopa table
@ 0        S1     1.0         3.                   0.                         1.0                 -0.9999                      0              12C-16O2_T400.aiopa
