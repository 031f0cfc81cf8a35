// Source compatibility with the reference tree layout (src/core/mp2v_hdr.h): the sequence-level header types.
#pragma once
#include "../mp2v_stream_headers.h"
