1e-2  1e2
1e-1  1e2
1e0   1e2
1e1   1e2
1e2   1e2
1e3   1e2
1e4   1e2
1e5   1e2
