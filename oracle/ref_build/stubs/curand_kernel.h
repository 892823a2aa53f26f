// stand-in: the reference only declares a dead `curandState s;` (kernel.cu:1439)
#pragma once
struct curandState { int unused; };
