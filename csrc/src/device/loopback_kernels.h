// Host twins of the collective / p2p kernels (loopback device backend)
#pragma once

#include "launch_api.h"

#include <vector>

namespace fb::host {

// true if (dtype, op) has an element-wise reduction
bool reducible(int dtype, int op);

bool waitFlagGe(const FbCommDev& c, const uint32_t* p, uint32_t target, uint32_t errCode);

int reduceKernel(const ReduceArgs& a, int dtype, int op, int blocks);
int llAllReduce(const LLArgs& a, int dtype, int op);
int groupAllReduce(const GroupArgs& a, int dtype, int op, int blocks); // a.segs: HOST memory
int moveKernel(const MoveArgs& a, int blocks);
int barrierKernel(const FbCommDev& c);
int p2pSend(const P2PArgs& a);
int p2pPull(const P2PArgs& a);
int putSignal(const PutArgs& a, int blocks);
int waitSignal(const FbCommDev& c, int signalIdx, uint32_t addTarget);
int signalPeers(const FbCommDev& c, uint32_t wordOff, uint32_t value);

}
