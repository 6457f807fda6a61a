/*
 * GpuVariantCaller.java — drop-in for VariantCaller (VariantCaller.java:42-239) behind the reference's plugin
 * seam: "derive from DenovoCaller, register in DenovoCallers" (README.rst:128-130, DenovoCallers.java:32-34).
 * SOURCE ONLY: the build image has no JDK, so this file has never been compiled; the compiled, tested twin is
 * the C++ GpuVariantCaller of host/denovo_host.hpp (same structure, same ABI calls).
 *
 * execute() keeps VariantCaller.execute's shape — contig list filtered by --chromosome, one position stride per
 * thread, a fresh join per stride — but a stride is no longer a VariantsBuffer fed call by call: its
 * (Variant, VariantCall) stream is flattened into dnv_vcall records (exactly the fields VariantsBuffer reads,
 * VariantsBuffer.java:172-306) and ONE downcall (dnv_mendel_filter) replaces the
 * checkAndAdd / canProcess / retrieveNextCall loop; PositionCalls come back in stream order.
 */
package com.google.cloud.genomics.denovo;

import com.google.api.services.genomics.model.ReferenceBound;
import com.google.api.services.genomics.model.Variant;
import com.google.api.services.genomics.model.VariantCall;
import com.google.cloud.genomics.denovo.DenovoUtil.Chromosome;
import com.google.cloud.genomics.denovo.DenovoUtil.Genotype;
import com.google.cloud.genomics.denovo.DenovoUtil.TrioMember;
import com.google.cloud.genomics.denovo.gpu.DenovoB200;

import java.io.File;
import java.io.IOException;
import java.io.PrintWriter;
import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.ValueLayout;
import java.util.ArrayList;
import java.util.List;
import java.util.concurrent.ExecutorService;
import java.util.concurrent.Executors;
import java.util.concurrent.TimeUnit;
import java.util.concurrent.atomic.AtomicInteger;

public class GpuVariantCaller extends DenovoCaller {

  private final DenovoShared shared;
  private final String libraryPath;
  private final AtomicInteger variantCounter = new AtomicInteger();

  public GpuVariantCaller(DenovoShared shared, String libraryPath) {
    this.shared = shared;
    this.libraryPath = libraryPath;
  }

  @Override
  public void execute() throws IOException {
    shared.getLogger().info("---- Starting Variant Caller ----");
    final File outputFile = DenovoUtil.getNormalizedFile(shared.getOutputFileName());
    List<ReferenceBound> allContigBounds = shared.getGenomics().variantsets().get(shared.getDatasetId())
        .setDisableGZipContent(true).execute().getReferenceBounds();
    try (PrintWriter callWriter = new PrintWriter(outputFile)) {
      ExecutorService executor = Executors.newFixedThreadPool(shared.getNumThreads());
      for (ReferenceBound contigBound : allContigBounds) {
        if (!shared.getChromosomes().contains(Chromosome.valueOf(contigBound.getReferenceName().toUpperCase()))) {
          continue; // VariantCaller.java:75-84
        }
        final String contig = contigBound.getReferenceName();
        long startContigPos = shared.getStartPosition() == null ? 1L : shared.getStartPosition();
        long endContigPos = shared.getEndPosition() == null ? contigBound.getUpperBound() : shared.getEndPosition();
        long strideLength = (endContigPos - startContigPos) / shared.getNumThreads(); // :92
        shared.getLogger().info("Processing Chromosome : " + contig);
        for (int threadIdx = 0; threadIdx < shared.getNumThreads(); threadIdx++) {
          final long start = startContigPos + threadIdx * strideLength;
          final long end = threadIdx == shared.getNumThreads() - 1 ? endContigPos : start + strideLength - 1;
          executor.execute(() -> {
            try {
              callSimpleDenovo(callWriter, contig, start, end);
            } catch (IOException e) { // VariantCaller.java:231-237: the shard is dropped, the run goes on
              e.printStackTrace();
              shared.getLogger().severe("Failed to run call for " + contig + ":" + start + "-" + end);
            }
          });
        }
      }
      executor.shutdown();
      try {
        executor.awaitTermination(Long.MAX_VALUE, TimeUnit.NANOSECONDS);
      } catch (InterruptedException e) {
        Thread.currentThread().interrupt();
      }
      shared.getLogger().info("---- Variant caller terminated ----");
    }
  }

  /** One position stride (VariantCaller.java:120-210): page through the stream, then ONE filter call. */
  void callSimpleDenovo(PrintWriter callWriter, String contig, Long startPosition, Long endPosition)
      throws IOException {
    VariantContigStream stream = new VariantContigStream(contig, startPosition, endPosition,
        new ArrayList<>(shared.getPersonToCallsetIdMap().values()), shared);
    List<Variant> variants = new ArrayList<>();
    while (stream.hasMore()) {
      List<Variant> page = stream.getVariants();
      if (page == null) {
        return; // VariantCaller.java:142-144: a null page ends the shard WITHOUT flushing
      }
      for (Variant v : page) {
        if (v.getCalls() == null) continue; // :147-153
        synchronized (this) {
          if (variantCounter.get() % DenovoUtil.VARIANT_LOG_FREQ == 0 && variantCounter.get() > 0) {
            shared.getLogger().info(String.format("%d Variant candidates processed", variantCounter.get()));
          }
          variantCounter.getAndIncrement();
        }
        variants.add(v);
      }
    }
    // a context per shard thread: dnv_ctx is not re-entrant (include/denovo_b200.h)
    try (DenovoB200 engine = new DenovoB200(libraryPath, shared.getSequenceErrorRate(),
        shared.getDenovoMutationRate(), shared.getLrtThreshold(), shared.getInferMethod().ordinal(), 0)) {
      callShard(engine, callWriter, contig, variants);
    }
  }

  /** Flatten one (variant, call) pair (see dnv_vcall in include/denovo_b200.h). */
  static void pack(MemorySegment rec, Variant variant, VariantCall call, TrioMember person, int variantId) {
    List<Integer> gt = call.getGenotype();
    String ref = variant.getReferenceBases();
    rec.set(ValueLayout.JAVA_LONG, 0, variant.getStart());
    rec.set(ValueLayout.JAVA_LONG, 8, variant.getEnd());
    rec.set(ValueLayout.JAVA_INT, 16, variantId);
    rec.set(ValueLayout.JAVA_BYTE, 20, (byte) (gt.size() > 0 ? gt.get(0) : 0));
    rec.set(ValueLayout.JAVA_BYTE, 21, (byte) (gt.size() > 1 ? gt.get(1) : 0));
    rec.set(ValueLayout.JAVA_BYTE, 22, (byte) Math.min(gt.size(), 255));
    rec.set(ValueLayout.JAVA_BYTE, 23, (byte) person.ordinal()); // CHILD 0, MOM 1, DAD 2
    List<String> filter = call.getInfo() == null ? null : call.getInfo().get("FILTER");
    boolean pass = filter != null && filter.size() == 1 && filter.get(0).equals("PASS"); // VariantsBuffer:245-249
    rec.set(ValueLayout.JAVA_BYTE, 24, (byte) (pass ? 1 : 0));
    rec.set(ValueLayout.JAVA_BYTE, 25, (byte) Math.min(ref.length(), 255));
    rec.set(ValueLayout.JAVA_BYTE, 26, (byte) (ref.isEmpty() ? 0 : ref.charAt(0)));
    for (int k = 0; k < 2; k++) {
      int g = gt.size() > k ? gt.get(k) : 0;
      String alt = g > 0 ? variant.getAlternateBases().get(g - 1) : "";
      rec.set(ValueLayout.JAVA_BYTE, 27 + 2L * k, (byte) Math.min(alt.length(), 255));
      rec.set(ValueLayout.JAVA_BYTE, 28 + 2L * k, (byte) (alt.isEmpty() ? 0 : alt.charAt(0)));
    }
  }

  /** The checkAndAdd / canProcess / retrieveNextCall loop of one start-sorted shard as ONE downcall. */
  void callShard(DenovoB200 engine, PrintWriter callWriter, String contig, List<Variant> variants) {
    long n = 0;
    for (Variant v : variants) n += v.getCalls().size();
    if (n == 0) return;
    try (Arena arena = Arena.ofConfined()) {
      MemorySegment recs = arena.allocate(n * DenovoB200.VCALL.byteSize(), 8);
      long i = 0;
      int vid = 0;
      for (Variant v : variants) {
        for (VariantCall call : v.getCalls()) {
          pack(recs.asSlice(i * 32, 32), v, call, shared.getCallsetIdToPersonMap().get(call.getCallSetId()), vid);
          i++;
        }
        vid++;
      }
      MemorySegment out = arena.allocate(n * DenovoB200.PCALL.byteSize(), 8);
      long nOut = engine.mendelFilter(recs, n, out, n);
      Genotype[] g = Genotype.values();
      StringBuilder sb = new StringBuilder();
      for (long k = 0; k < nOut; k++) {
        MemorySegment c = out.asSlice(k * 24, 24);
        if (c.get(ValueLayout.JAVA_BYTE, 19) == 0) continue; // is_denovo (VariantCaller.java:176-183)
        long pos = c.get(ValueLayout.JAVA_LONG, 0);
        // "%s,%d,%s%n" with PositionCall.toString (VariantCaller.java:178-179, VariantsBuffer.java:358-360)
        String line = String.format("%s,%d,[%d,{CHILD=%s, MOM=%s, DAD=%s}]", contig, pos, pos,
            g[c.get(ValueLayout.JAVA_BYTE, 16)], g[c.get(ValueLayout.JAVA_BYTE, 17)], g[c.get(ValueLayout.JAVA_BYTE, 18)]);
        shared.getLogger().fine(line);
        sb.append(line).append(String.format("%n"));
      }
      writeCalls(callWriter, sb.toString());
    }
  }
}
