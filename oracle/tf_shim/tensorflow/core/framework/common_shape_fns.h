// Test-infrastructure shim: forwards to the single header that stands in for the
// TensorFlow framework surface used by the reference Conv3p CPU op.
#pragma once
#include "../../../tf_shim.h"
