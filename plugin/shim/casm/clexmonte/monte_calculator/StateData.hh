// forwards to the stand-in declarations (plugin/shim/casm_shim.hh); with libcasm installed the real header is found instead
#pragma once
#include "casm_shim.hh"
