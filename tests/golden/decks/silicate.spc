#
# This file contains a template for a list of species in the form of a table
# every row is one species and starts with a * and a number. Row contents are listed below
# 
# 1:Number 2:Name     3:mass in amu 4:molar dof        5:adiabatic index initially 6:relative amount 7:initial density excess/P2, velocity/P1 8: is_dust_like?  9:opacity.opa  10:opa file length
@ 0        Gas        30.           7.                 1.29                        1.0                 1.                                     0                 smalldust.opa  8
@ 1        Grain0.1   30.           2.523e11           1.01                        0.007               1.                                     1                 smalldust.opa  8
