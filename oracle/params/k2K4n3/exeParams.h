#ifndef _EXE_PARAMS_H_
#define _EXE_PARAMS_H_
#define ENCODE_BLOCK_SIZE 64
#define DECODE_BLOCK_SIZE 64
#endif
