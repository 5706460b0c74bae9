#
# This file contains a template for a list of species in the form of a table
# every row is one species and starts with a * and a number. Row contents are listed below
# 
# 1:Number 2:Name 3:mass in amu 4:molar dof        5:adiabatic index initially 6:relative amount 7:initial density excess/P2, velocity/P1 8: is_dust_like?  9:opacity.opa 10:opa file length
@ 0        H0     2.            5.                    1.4                         1                 -1.                                      0                smalldust.opa  8
@ 1        H1     2.            5.                    1.4                         1.5                2.                                      0                smalldust.opa  8
@ 2        H2     2.            5.                    1.4                         2.0                3.1                                     0                smalldust.opa  8
@ 3        H3     2.            5.                    1.4                         2.5                -2.5                                    0                smalldust.opa  8
@ 4        H4     2.            5.                    1.4                         3.0                0.5                                     0                smalldust.opa  8
@ 5        H5     2.            5.                    1.4                         3.5                -4.1                                    0                smalldust.opa  8
