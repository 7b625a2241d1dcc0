// Same declarations as the reference's dist/tsc/zlib.d.ts, plus the batch forms.
export declare function inflate(input: Uint8Array): Uint8Array;
export declare function deflate(input: Uint8Array): Uint8Array;
export declare function inflateBatch(inputs: Uint8Array[]): Uint8Array[];
export declare function deflateBatch(inputs: Uint8Array[]): Uint8Array[];
export declare function deflateRaw(input: Uint8Array): Uint8Array;
export declare function inflateRaw(input: Uint8Array): Uint8Array;
export declare function gzip(input: Uint8Array): Uint8Array;
export declare function gunzip(input: Uint8Array): Uint8Array;
