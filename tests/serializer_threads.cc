// TEST INFRASTRUCTURE (tests/test_serializer_host.py): the frame writer's parallel sections under ThreadSanitizer.
// Four threads each write 1080p frames through a HostPool group while the loop-filter level arrives late from the
// submitting thread (EncodeFeatures::late_loop_filter_level, the way encoder.cu encode_final runs the writer next to
// the device's loop-filter search); the pool's helping waits, the chunked recording of a one-partition frame and the
// concurrent token / first partition writers all run.  Key frames must equal the single-threaded result byte for byte.
// usage: serializer_threads FILE.ivf   (exit code 0 = same bytes; TSan reports go to stderr and fail the run)
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
#include "hostpool.h"
#include "parser.h"
#include "serializer.h"
using namespace vp8;
struct Late { std::mutex m; std::condition_variable cv; bool ready=false; int level=0; };
static int wait_level(void* p){ Late* l=(Late*)p; std::unique_lock<std::mutex> lk(l->m); l->cv.wait(lk,[l]{return l->ready;}); return l->level; }
int main(int argc,char**argv){
  FILE*f=fopen(argv[1],"rb"); std::vector<uint8_t> d; uint8_t buf[65536]; size_t n; while((n=fread(buf,1,sizeof buf,f))>0) d.insert(d.end(),buf,buf+n); fclose(f);
  int w=d[12]|d[13]<<8,h=d[14]|d[15]<<8; std::vector<std::pair<size_t,size_t>> fr; size_t p=32; while(p+12<=d.size()){ uint32_t sz; memcpy(&sz,&d[p],4); fr.push_back({p+12,sz}); p+=12+sz; }
  State st(w,h); std::vector<ParsedFrame*> pfs;
  for(int i=0;i<3;i++){ ParsedFrame* pf=new ParsedFrame(); if(parse_frame(st,&d[fr[i].first],fr[i].second,*pf)) return 1; pfs.push_back(pf); }
  std::vector<std::vector<uint8_t>> ref(3);
  for(int i=0;i<3;i++){ EncodeHeader hd; hd.key_frame=pfs[i]->desc.key_frame; hd.show_frame=true; hd.width=w; hd.height=h; hd.y_ac_qi=40; hd.loop_filter_level=7+i; hd.sharpness=0; hd.optimize_token_probs=true;
    ref[i]=serialize_frame(hd,pfs[i]->mbs.data(),pfs[i]->tokens.data(),pfs[i]->split.data(),nullptr); if(ref[i].empty()) {printf("serialize failed\n"); return 1;} }
  std::atomic<int> bad{0};
  std::vector<std::thread> th;
  for(int t=0;t<4;t++) th.emplace_back([&,t]{
    for(int it=0;it<6;it++){ int i=(t+it)%3; EncodeHeader hd; hd.key_frame=pfs[i]->desc.key_frame; hd.show_frame=true; hd.width=w; hd.height=h; hd.y_ac_qi=40; hd.loop_filter_level=0; hd.sharpness=0; hd.optimize_token_probs=true;
      EncodeFeatures ft; Late late; ft.late_loop_filter_level=&wait_level; ft.late_ctx=&late; std::vector<uint8_t> out;
      HostPool::Group g; g.run([&]{ out=serialize_frame(hd,pfs[i]->mbs.data(),pfs[i]->tokens.data(),pfs[i]->split.data(),&ft); });
      std::this_thread::sleep_for(std::chrono::microseconds(300*(it%3)));
      { std::lock_guard<std::mutex> lk(late.m); late.level=7+i; late.ready=true; } late.cv.notify_all();
      g.wait();
      // features != nullptr changes prob_last/golden defaults for inter frames, so compare only key frames byte for byte; others by size sanity
      if(pfs[i]->desc.key_frame ? out!=ref[i] : out.empty()) bad++;
    }});
  for(auto&t:th)t.join();
  printf("bad %d\n",bad.load()); return bad.load()!=0;
}
