// code_tables_dev.cuh -- the __constant__ copy of the QC description (struct CodeTables, decode_kernels.cuh), used by the
// table-driven kernels (finalize, encoder).  Defined in the main translation unit only: the decode kernels take block
// columns and shifts as literals and are built in their own translation units (decode_inst.cu).
#pragma once
#include "decode_kernels.cuh"

namespace ldpc {
__constant__ CodeTables c_code;
}
