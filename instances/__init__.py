"""Host-side builders of the benchmark / test INSTANCES (SeDuMi-format inputs of BASELINE.json's configs): restatements
of the reference's problem generators src/basicfunction/{get_basis,bqpmom,qsmom,Laplacian}.m and
example/generate_hamming.m.  Input generation is neither the product (manisdp_matlab_b200/) nor the checker
(oracle/): tests, tools and bench.py import it to build the inputs both of them are then run on."""
