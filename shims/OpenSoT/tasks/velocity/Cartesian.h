// Oracle build shim: AutoStack.h includes this model-dependent header; the solver path never uses it.
#pragma once
