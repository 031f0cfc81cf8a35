// Source-compatibility shim: code written against the reference's "core/decoder.h"
// (fxslava/tiny_mp2v_dec src/core/decoder.h:25-131) compiles against this repository unchanged.
#pragma once
#include "common/cpu.hpp"
#include "../mp2v_decoder.hpp"
