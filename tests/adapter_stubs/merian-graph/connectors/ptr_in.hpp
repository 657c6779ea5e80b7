// stub: see tests/adapter_stubs/merian_stub.hpp
#pragma once
#include "merian_stub.hpp"
