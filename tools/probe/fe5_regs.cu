// Dev probe: compile nww_fe5.cuh alone to read its register / spill report (nvcc -Xptxas -v), no host code.
#include "../../nanowakeword_b200/csrc/nww_fe5.cuh"
