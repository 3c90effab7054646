/*
 * Extension (not in the reference): the reference's decode loop -- one VITERBI_DECODER_HARD(..., last=true)
 * call per packet, speedDecode/speedDecode.c:78-79, berTestK7/berTestK7.c:150 -- handed to the GPU a batch at a
 * time (SURVEY 8(f)1 "batching of consecutive last=true calls").  A synchronous one-packet call can only ever
 * use one warp of the GPU; with the queue the loop keeps its shape and two lines change:
 *
 *     VITERBI_DECODER_HARD(&state, coded[i], decoded[i], segs, true);    ->   viterbiQueueSubmit(q, coded[i], decoded[i]);
 *     (after the loop)                                                   ->   viterbiQueueFlush(q);
 *
 * `decoded[i]` is written some time between its Submit and the return of the next Flush.  Packets of one queue
 * all have the same length; the code parameters are the library's (convCodeParams.h).  One producer thread per
 * queue.  Failures follow the reference's convention: message on stdout and exit(1).
 */
#ifndef _H_VITERBI_DECODER_QUEUE
#define _H_VITERBI_DECODER_QUEUE

#include <stdint.h>

typedef struct viterbiQueue viterbiQueue_t;

/* packets of `segmentsPerPacket` coded segments (8*bytes/k + S, convEncode.h:62); `maxPackets` are collected
 * before a batch is handed to the GPU (two batches exist: one filling, one decoding) */
viterbiQueue_t *viterbiQueueCreate(int segmentsPerPacket, int maxPackets);
/* copies the packet; returns the number of bytes that will be written to `uncoded` */
int viterbiQueueSubmit(viterbiQueue_t *q, const uint8_t *codedSegments, uint8_t *uncoded);
/* waits until every submitted packet is decoded and delivered; returns how many were delivered since the last flush */
long viterbiQueueFlush(viterbiQueue_t *q);
void viterbiQueueDestroy(viterbiQueue_t *q);

/* The same for the encoder loop (speedEncode/speedEncode.c:65-67, one convEnc(..., last=true) per packet):
 * `codedSegments` receives 8*bytesPerPacket/k + S segments some time before the next Flush returns. */
typedef struct viterbiQueue convEncQueue_t;
convEncQueue_t *convEncQueueCreate(int bytesPerPacket, int maxPackets);
int convEncQueueSubmit(convEncQueue_t *q, const uint8_t *uncoded, uint8_t *codedSegments);
long convEncQueueFlush(convEncQueue_t *q);
void convEncQueueDestroy(convEncQueue_t *q);

#endif
