#
# This file contains a template for a list of species in the form of a table
# every row is one species and starts with a * and a number. Row contents are listed below
# 
# 1:Number 2:Name 3:mass in amu 4:dof               5:adiabatic index initially 6:relative amount 7:initial density excess 8: is_dust_like 9:opacity.opa 10:opa file length
@ 0        S1     1.0         3.                   0.                         1.0                 -0.9999                      0              highenergy3.opa  
@ 1        S2     16.0        3.                   0.                         1.e-10                 -0.0                      0              highenergy_protons.opa  
@ 2        S3     5e-4        3.                   0.                         1.e-10                 -0.0                      0             highenergy_electrons.opa  
