#
# This file contains a template for a list of species in the form of a table
# every row is one species and starts with a * and a number. Row contents are listed below
# 
# 1:Number 2:Name 3:mass in amu 4:dof               5:adiabatic index initially 6:relative amount 7:initial density excess 8: is_dust_like 9:opacity.opa 10:opa file length
@ 0        H0     0.6          21.                   1.4                         1.e0                 0.                      0           largedust.opa  8
@ 1        H1     0.6          21.                   1.4                         1.e-5              0.                      0              largedust.opa  8
@ 2        H2     0.6          21.                   1.4                         1.e-5              0.                      0              largedust.opa  8
@ 3        H3     0.6          21.                   1.4                         1.e-5              0.                      0              largedust.opa  8
@ 4        H4     0.6          21.                   1.4                         1.e-5              0.                      0              largedust.opa  8
@ 5        H5     0.6          21.                   1.4                         1.e-5              0.                      0              largedust.opa  8
@ 6        H6     0.6          21.                   1.4                         1.e-5              0.                      0              largedust.opa  8
@ 7        H7     0.6          21.                   1.4                         1.e-5              0.                      0              largedust.opa  8
@ 8        H8     0.6          21.                   1.4                         1.e-5              0.                      0              largedust.opa  8
@ 9        H9     0.6          21.                   1.4                         1.e-5              0.                      0              largedust.opa  8
@ 10       H10     0.6          21.                   1.4                         1.e-5              0.                      0              largedust.opa  8
@ 11       H11     0.6          21.                   1.4                         1.e-5              0.                      0              largedust.opa  8
@ 12       H12     0.6          21.                   1.4                         1.e-5              0.                      0              largedust.opa  8
