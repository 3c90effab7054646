/* Block-size knobs kept for source compatibility with the reference's
 * src/defaultParams/exeParams.h:4-6 (nothing in either code base reads them). */
#ifndef CED_PARAMS_DEFAULT_EXE_PARAMS_H
#define CED_PARAMS_DEFAULT_EXE_PARAMS_H
#define DECODE_BLOCK_SIZE 64
#define ENCODE_BLOCK_SIZE 64
#endif
